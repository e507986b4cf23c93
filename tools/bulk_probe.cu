// Developer microbenchmark: latency and throughput of cp.async.bulk global->shared (L2-resident source).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// depth = copies in flight; each copy `bytes`; n copies total per CTA; every CTA reads the same 1.2 MB region (like the weights)
__global__ void __launch_bounds__(128, 1) probe(const uint8_t* src, int bytes, int depth, int n, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[16];
    if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&bars[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t sb = smem_u32(smem);
        uint32_t phase[16] = {0};
        long long t0 = clock64();
        long long lat = 0;
        for (int c = 0; c < n + depth; ++c) {
            int s = c % depth;
            if (c >= depth) { while (!mbar_try_wait(smem_u32(&bars[s]), phase[s])) {} phase[s] ^= 1; }
            if (c < n) {
                mbar_expect_tx(smem_u32(&bars[s]), bytes);
                bulk_g2s(sb + s * bytes, src + (size_t)((c * 16384) % (1200 * 1024)), bytes, smem_u32(&bars[s]));
            }
            if (depth == 1 && c < n) { long long a = clock64(); while (!mbar_try_wait(smem_u32(&bars[0]), phase[0])) {} lat += clock64() - a; phase[0] ^= 1; c += 0; mbar_expect_tx(smem_u32(&bars[0]), 0); phase[0] ^= 0; }
        }
        long long t1 = clock64();
        out[blockIdx.x * 2] = t1 - t0;
        out[blockIdx.x * 2 + 1] = lat;
    }
}
__global__ void __launch_bounds__(128, 1) latency(const uint8_t* src, int bytes, int n, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t ph = 0; long long tot = 0;
        for (int c = 0; c < n; ++c) {
            long long a = clock64();
            mbar_expect_tx(smem_u32(&bar), bytes);
            bulk_g2s(smem_u32(smem), src + (size_t)((c * 16384) % (1200 * 1024)), bytes, smem_u32(&bar));
            while (!mbar_try_wait(smem_u32(&bar), ph)) {}
            ph ^= 1;
            tot += clock64() - a;
        }
        out[blockIdx.x] = tot;
    }
}
int main() {
    uint8_t* g; cudaMalloc(&g, 4 << 20); cudaMemset(g, 1, 4 << 20);
    long long* out; cudaMalloc(&out, 148 * 2 * sizeof(long long));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(latency, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long h[296];
    for (int grid : {1, 148}) for (int bytes : {1024, 8192, 16384, 32768}) {
        latency<<<grid, 128, 200 * 1024>>>(g, bytes, 200, out); cudaDeviceSynchronize();
        cudaMemcpy(h, out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
        double mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("latency grid=%3d bytes=%5d : %.0f cycles per copy (serialised)\n", grid, bytes, mx / 200);
    }
    for (int grid : {1, 148}) for (int bytes : {8192, 16384}) for (int depth : {2, 3, 4, 8}) {
        if (bytes * depth > 196 * 1024) continue;
        int n = 2000;
        probe<<<grid, 128, 200 * 1024>>>(g, bytes, depth, n, out); cudaDeviceSynchronize();
        cudaMemcpy(h, out, sizeof(long long) * grid * 2, cudaMemcpyDeviceToHost);
        double mx = 0; for (int i = 0; i < grid; ++i) mx = h[2 * i] > mx ? h[2 * i] : mx;
        printf("stream  grid=%3d bytes=%5d depth=%d : %.1f B/cycle/SM  (%.0f cycles per copy)\n", grid, bytes, depth, (double)bytes * n / mx, mx / n);
    }
    return 0;
}
