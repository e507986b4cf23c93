// Developer probe: what throttles a layer epilogue (TMEM -> regs -> bias/convert -> swizzled smem) while the tensor
// pipe of the same SM is busy? 4 or 8 epilogue warps drain 128 x 256 fp32 accumulator passes from TMEM columns
// 0..255, with/without the conversion math and the st.shared, with/without a concurrent tcgen05.mma stream
// (M=128, N=256, K=16, operands in other shared memory, accumulating into columns 256..511).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/epi_probe tools/epi_probe.cu && tools/epi_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) { return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61); }
__host__ __device__ constexpr uint32_t umma_idesc(int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ void tmem_ld32(uint32_t a, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
          "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]) : "r"(a));
}
__device__ __forceinline__ void tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// smem: [0,64K) epilogue output, [64K,128K) A operand, [128K,192K) B operand, then bias
struct BiasArg { float v[11][256]; };
template <int mode>
__global__ void __launch_bounds__(320, 1) probe(const __grid_constant__ BiasArg ba, int epi_warps, int with_mma, int passes, long long* out, uint32_t* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tmem_ptr;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 9) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    for (int i = threadIdx.x; i < 194 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_ptr, sb = smem_u32(smem);
    const float* s_bias = reinterpret_cast<const float*>(smem + 192 * 1024);
    if (warp == 9) {
        if (with_mma && lane == 0) {
            const uint32_t idesc = umma_idesc(256);
            long long n = 0;
            while (!stop) {
                const uint64_t ad = umma_desc(sb + 65536u + (uint32_t)(n & 3) * 16384u), bd = umma_desc(sb + 131072u + (uint32_t)(n & 1) * 32768u);
                // at most two batches of 4 MMAs in flight: batch n waits for batch n - 2 (same barrier, previous phase)
                if (n >= 2) while (!mbar_try_wait(smem_u32(&bars[n & 1]), (uint32_t)(((n >> 1) - 1) & 1))) {}
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(tm + 256u, ad + 2 * k, bd + 2 * k, idesc, 1u);
                umma_commit(smem_u32(&bars[n & 1]));
                ++n;
            }
            out[blockIdx.x * 16 + 15] = n * 4;
        }
    } else if (warp < epi_warps) {
        const int q = warp & 3, h = warp >> 2, row = q * 32 + lane;
        const int ngroups = epi_warps == 8 ? 4 : 8;
        const uint32_t tbase = tm + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 128);
        uint32_t accx = 0;
        long long t0 = clock64();
        for (int p = 0; p < passes; ++p) {
            uint32_t r[2][32];
            tmem_ld32(tbase, r[0]);
            float4 bq[8];
            if (mode & 64) {
#pragma unroll
                for (int i = 0; i < 8; ++i) bq[i] = reinterpret_cast<const float4*>(s_bias)[i];
            }
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                if (g >= ngroups) break;
                uint32_t (&rr)[32] = r[g & 1];
                tmem_wait();
                if (g + 1 < ngroups) tmem_ld32(tbase + 32u * (g + 1), r[(g + 1) & 1]);
                if (mode == 0) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) accx ^= rr[i];
                } else {
                    // mode bits: 1 = math on, 2 = bias from smem (else a constant), 4 = integer conversion for odd pairs,
                    //            8 = integer conversion for all pairs, 16 = st.shared, 32 = skip the bias add
                    const float4* b4 = reinterpret_cast<const float4*>(s_bias + 32 * g);
                    uint32_t o[16];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float4 bb;
                        if (mode & 128) {           // bias from the kernel parameters (constant bank, warp-uniform address)
                            const float* bp = ba.v[p % 11] + 32 * g + 4 * i;
                            bb = make_float4(bp[0], bp[1], bp[2], bp[3]);
                        } else if (mode & 64) { bb = bq[i]; if (g + 1 < ngroups) bq[i] = b4[8 + i]; }
                        else bb = (mode & 2) ? b4[i] : make_float4(0.5f, 0.25f, 0.125f, 1.f);
                        float2 s0 = make_float2(__uint_as_float(rr[4 * i]), __uint_as_float(rr[4 * i + 1]));
                        float2 s1 = make_float2(__uint_as_float(rr[4 * i + 2]), __uint_as_float(rr[4 * i + 3]));
                        if (!(mode & 32)) { s0 = __fadd2_rn(s0, make_float2(bb.x, bb.y)); s1 = __fadd2_rn(s1, make_float2(bb.z, bb.w)); }
                        if (mode & 8) {
                            int a0 = max(__float_as_int(s0.x), 0), a1 = max(__float_as_int(s0.y), 0);
                            o[2 * i] = __byte_perm((uint32_t)a0 + 0x8000u, (uint32_t)a1 + 0x8000u, 0x7632);
                        } else {
                            asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(o[2 * i]) : "f"(s0.y), "f"(s0.x));
                        }
                        if (mode & 12) {
                            int a0 = max(__float_as_int(s1.x), 0), a1 = max(__float_as_int(s1.y), 0);
                            o[2 * i + 1] = __byte_perm((uint32_t)a0 + 0x8000u, (uint32_t)a1 + 0x8000u, 0x7632);
                        } else {
                            asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(o[2 * i + 1]) : "f"(s1.y), "f"(s1.x));
                        }
                    }
                    if (!(mode & 16)) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) accx ^= o[i];
                    } else {
                        uint8_t* chunk = smem + ((h * 4 + g) >> 1) * 16384;
                        const int u0 = (g & 1) * 4;
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            *reinterpret_cast<uint4*>(chunk + row * 128 + (((u0 + u) ^ (row & 7)) << 4)) = make_uint4(o[4 * u], o[4 * u + 1], o[4 * u + 2], o[4 * u + 3]);
                    }
                }
            }
        }
        long long t1 = clock64();
        if (lane == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
        if (accx == 0x12345678u) sink[0] = accx;
    }
    // epilogue warps done -> stop the MMA stream
    if (warp < epi_warps) asm volatile("bar.sync 1, %0;" ::"r"(epi_warps * 32));
    if (threadIdx.x == 0) stop = 1;
    __syncthreads();
    if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

template <int M>
static void go(int epi_warps, int with_mma, int passes, long long* out, uint32_t* sink, int smem) {
    cudaFuncSetAttribute(probe<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    static BiasArg ba;
    for (int j = 0; j < 11; ++j) for (int c = 0; c < 256; ++c) ba.v[j][c] = 0.001f * (c + j);
    probe<M><<<148, 320, smem>>>(ba, epi_warps, with_mma, passes, out, sink);
}
static void launch(int m, int epi_warps, int with_mma, int passes, long long* out, uint32_t* sink, int smem) {
    switch (m) {
        case 0: go<0>(epi_warps, with_mma, passes, out, sink, smem); break;
        case 1: go<(1 | 2 | 16)>(epi_warps, with_mma, passes, out, sink, smem); break;
        case 2: go<(1 | 16)>(epi_warps, with_mma, passes, out, sink, smem); break;
        case 3: go<(1 | 2 | 4 | 16)>(epi_warps, with_mma, passes, out, sink, smem); break;
        case 4: go<(1 | 2 | 8 | 16)>(epi_warps, with_mma, passes, out, sink, smem); break;
        case 5: go<(1 | 2 | 16 | 32)>(epi_warps, with_mma, passes, out, sink, smem); break;
        case 6: go<(1 | 8 | 16 | 32)>(epi_warps, with_mma, passes, out, sink, smem); break;
        case 7: go<(1 | 2)>(epi_warps, with_mma, passes, out, sink, smem); break;
        case 8: go<(1 | 2 | 16 | 64)>(epi_warps, with_mma, passes, out, sink, smem); break;
        default: go<(1 | 16 | 128)>(epi_warps, with_mma, passes, out, sink, smem); break;
    }
}
int main() {
    long long* out; cudaMalloc(&out, 148 * 16 * sizeof(long long)); uint32_t* sink; cudaMalloc(&sink, 4);
    const int smem = 196 * 1024;
    const int passes = 200;
    const char* names[] = {"ld only", "full (bias smem, F2FP, STS)", "bias const", "mixed F2FP/int", "all int", "no bias add", "int, no bias add", "full, no STS", "full, bias prefetched 1 group", "full, bias from kernel params"};
    for (int epi_warps : {4, 8})
        for (int m = 0; m < 10; ++m)
            for (int with_mma = 0; with_mma < 2; ++with_mma) {
                cudaMemset(out, 0, 148 * 16 * sizeof(long long));
                launch(m, epi_warps, with_mma, passes, out, sink, smem);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[16]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                long long mx = 0; for (int i = 0; i < epi_warps; ++i) mx = h[i] > mx ? h[i] : mx;
                printf("warps=%d %-30s mma=%d : %6.0f cycles / 128x256 pass   mma %.0f cyc each %s\n", epi_warps, names[m], with_mma,
                       (double)mx / passes, h[15] ? (double)mx / h[15] : 0.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
