// Developer microbenchmark (not part of the product): cycles per tcgen05.mma for a few shapes, issued
// back to back by one thread on fixed shared-memory operands, optionally with a concurrent bulk-copy
// stream into shared memory. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) { return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61); }
__host__ __device__ constexpr uint32_t umma_idesc(int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24); }

// mode bit0: concurrent bulk copies (16 KB each, 3 in flight); N: MMA N; per_commit: MMAs per commit
__global__ void __launch_bounds__(128, 1) probe(int N, int iters, int per_commit, int with_copy, int mode_elect, const uint8_t* gsrc, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[8];
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    // zero operands
    for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_ptr;
    const uint32_t sb = smem_u32(smem);
    if (warp == 1 && (mode_elect || lane == 0)) {
        const uint32_t idesc = umma_idesc(N);
        long long t0 = clock64();
        for (int i = 0; i < iters; i += 4) {
            // A: 64 KB region (4 chunks), B: next 64 KB; 4 K-steps per chunk like the real kernel
            uint32_t a = sb + (uint32_t)((i >> 2) & 3) * 16384u;
            uint32_t b = sb + 65536u + (uint32_t)((i >> 2) & 1) * 32768u;
            const uint64_t ad = umma_desc(a), bd = umma_desc(b);
            const uint32_t d = tm + (uint32_t)((i & 4) ? 256 : 0);
            if (!mode_elect || elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(d, ad + 2 * k, bd + 2 * k, idesc, 1u);
                if (per_commit == 4) umma_commit(smem_u32(&bars[0]));
            }
            if (mode_elect) __syncwarp();
        }
        long long t1 = clock64();
        if (!mode_elect || elect_one()) umma_commit(smem_u32(&bars[1]));
        mbar_wait(smem_u32(&bars[1]), 0);
        long long t2 = clock64();
        if (lane == 0) { out[blockIdx.x * 4 + 0] = t1 - t0; out[blockIdx.x * 4 + 1] = t2 - t0; }
    } else if (warp == 2 && lane == 0 && with_copy) {
        // stream 16 KB chunks into a 3-deep ring beyond the operands (128K..176K)
        uint32_t phase[3] = {0, 0, 0};
        int n = iters / 4 * (N == 256 ? 2 : 1);   // same bytes per MMA work as the real kernel (16 KB per 4 N=128 MMAs)
        for (int c = 0; c < n; ++c) {
            int s = c % 3;
            if (c >= 3) { mbar_wait(smem_u32(&bars[2 + s]), phase[s]); phase[s] ^= 1; }
            mbar_expect_tx(smem_u32(&bars[2 + s]), 16384);
            bulk_g2s(sb + 131072u + s * 16384u, gsrc + (size_t)(c % 64) * 16384, 16384, smem_u32(&bars[2 + s]));
        }
        for (int s = 0; s < 3; ++s) if (n > s) mbar_wait(smem_u32(&bars[2 + s]), phase[s]);
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

int main() {
    uint8_t* g; cudaMalloc(&g, 64 * 16384); cudaMemset(g, 0, 64 * 16384);
    long long* out; cudaMalloc(&out, 148 * 4 * sizeof(long long));
    int smem = 200 * 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 4096;
    for (int grid : {148})
        for (int N : {64, 128, 256})
            for (int per_commit : {4, 1024})
                for (int with_copy : {0, 1}) for (int mode_elect : {0, 1}) {
                    probe<<<grid, 128, smem>>>(N, iters, per_commit, with_copy, mode_elect, g, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    long long h[4]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                    printf("grid=%3d N=%3d commit/%-4d copy=%d elect=%d : issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %d)  %s\n", grid, N, per_commit,
                           with_copy, mode_elect, (double)h[0] / iters, (double)h[1] / iters, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
                }
    return 0;
}
