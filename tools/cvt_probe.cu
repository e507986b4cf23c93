// Developer probe: throughput of fp32 -> bf16x2 conversion paths on one SM (results per clock per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/cvt_probe tools/cvt_probe.cu && tools/cvt_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(const float* in, uint32_t* out, int iters, long long* cycles) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = in[threadIdx.x * 16 + i];
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            uint32_t d;
            if (MODE == 0) {            // F2FP pack
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(v[i + 1]), "f"(v[i]));
            } else if (MODE == 1) {     // F2FP pack with relu
                asm volatile("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(v[i + 1]), "f"(v[i]));
            } else if (MODE == 2) {     // integer round-to-nearest-even + PRMT
                uint32_t a = __float_as_uint(v[i]), b = __float_as_uint(v[i + 1]);
                a += 0x7FFFu + ((a >> 16) & 1u);
                b += 0x7FFFu + ((b >> 16) & 1u);
                d = __byte_perm(a, b, 0x7632);
                asm volatile("" : "+r"(d));
            } else if (MODE == 3) {     // fp16 pack
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(v[i + 1]), "f"(v[i]));
            } else {                    // truncation only (PRMT), the lower bound on ALU cost
                d = __byte_perm(__float_as_uint(v[i]), __float_as_uint(v[i + 1]), 0x7632);
                asm volatile("" : "+r"(d));
            }
            acc ^= d;
            v[i] = __uint_as_float(__float_as_uint(v[i]) ^ (acc & 1u));   // keep a dependency so nothing is hoisted
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    float* in; uint32_t* out; long long* cyc;
    cudaMalloc(&in, 1024 * 16 * 4); cudaMalloc(&out, 1024 * 4 * 4); cudaMalloc(&cyc, 64);
    cudaMemset(in, 0x3f, 1024 * 16 * 4);
    const int iters = 2000;
    const char* names[5] = {"cvt.rn.bf16x2.f32", "cvt.rn.relu.bf16x2.f32", "integer RNE + PRMT", "cvt.rn.f16x2.f32", "PRMT truncation"};
    for (int threads : {128, 256, 512, 1024}) {
        for (int mode = 0; mode < 5; ++mode) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) probe<0><<<1, threads>>>(in, out, iters, cyc);
                if (mode == 1) probe<1><<<1, threads>>>(in, out, iters, cyc);
                if (mode == 2) probe<2><<<1, threads>>>(in, out, iters, cyc);
                if (mode == 3) probe<3><<<1, threads>>>(in, out, iters, cyc);
                if (mode == 4) probe<4><<<1, threads>>>(in, out, iters, cyc);
                cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            }
            double results = (double)threads * iters * 16;
            printf("threads %4d  %-24s %9lld cycles  %6.1f fp32 values converted / clk / SM\n", threads, names[mode], h, results / (double)h);
        }
    }
    return 0;
}
