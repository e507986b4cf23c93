// Developer probe: can the layer bias be added by the tensor core? One K=16 tcgen05.mma step with
//   A = "ones" operand: K-major, NO swizzle, two 8x16B core matrices (k 0..7: [1,1,0,...], k 8..15: zeros), SBO = 0 so that
//       all sixteen 8-row groups alias the same 256 bytes,
//   B = bias tile: K-major, NO swizzle, one 8x16B core matrix per 8 N-rows (row n: [hi(n), lo(n), 0...]), LBO = 0 so that
//       k 8..15 alias k 0..7 (they meet A's zeros),
// must give D[m][n] = hi(n) + lo(n) for every row m. Prints the max error against fp32 bias values.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bias_mma_probe tools/bias_mma_probe.cu && tools/bias_mma_probe
#include <cstdio>
#include <cstdint>
#include <cmath>
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
// no-swizzle K-major descriptor: LBO = bytes between core matrices along K, SBO = bytes between 8-row groups
__device__ __forceinline__ uint64_t desc_noswz(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t umma_idesc(int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24); }

__global__ void __launch_bounds__(128, 1) probe(const float* bias, float* out, int lbo_a, int sbo_a, int lbo_b, int sbo_b) {
    __shared__ __align__(1024) uint8_t ones[256];
    __shared__ __align__(1024) uint8_t btile[256 * 16];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(256)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    // ones: core matrix 0 = 8 rows x [1,1,0,0,0,0,0,0], core matrix 1 = zeros
    for (int i = threadIdx.x; i < 64; i += blockDim.x) {
        const int row = i / 8, k = i % 8;
        reinterpret_cast<__nv_bfloat16*>(ones)[row * 8 + k] = __float2bfloat16(k < 2 ? 1.f : 0.f);
        reinterpret_cast<__nv_bfloat16*>(ones)[64 + i] = __float2bfloat16(0.f);
    }
    // bias tile: N-row n at byte n*16: [hi, lo, 0, 0, 0, 0, 0, 0]
    for (int n = threadIdx.x; n < 256; n += blockDim.x) {
        const float b = bias[n];
        const __nv_bfloat16 hi = __float2bfloat16(b);
        const __nv_bfloat16 lo = __float2bfloat16(b - __bfloat162float(hi));
        __nv_bfloat16* r = reinterpret_cast<__nv_bfloat16*>(btile + n * 16);
        r[0] = hi; r[1] = lo;
        for (int k = 2; k < 8; ++k) r[k] = __float2bfloat16(0.f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_ptr;
    if (threadIdx.x == 0) {
        umma_f16(tm, desc_noswz(smem_u32(ones), lbo_a, sbo_a), desc_noswz(smem_u32(btile), lbo_b, sbo_b), umma_idesc(256), 0u);
        umma_commit(smem_u32(&bar));
    }
    while (!mbar_try_wait(smem_u32(&bar), 0)) {}
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // row = thread; read 256 columns
    for (int c0 = 0; c0 < 256; c0 += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) out[(size_t)(warp * 32 + lane) * 256 + c0 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256));
}

int main() {
    float hb[256], *db, *dout;
    for (int n = 0; n < 256; ++n) hb[n] = 0.37f * sinf(1.7f * n) + 1e-3f * n;
    cudaMalloc(&db, sizeof(hb)); cudaMalloc(&dout, 128 * 256 * 4);
    cudaMemcpy(db, hb, sizeof(hb), cudaMemcpyHostToDevice);
    static float ho[128 * 256];
    const int cfg[][4] = {{128, 0, 0, 128}, {0, 128, 128, 0}, {128, 0, 128, 0}, {0, 128, 0, 128}};   // lbo_a, sbo_a, lbo_b, sbo_b
    for (auto& c : cfg) {
        cudaMemset(dout, 0xff, 128 * 256 * 4);
        probe<<<1, 128>>>(db, dout, c[0], c[1], c[2], c[3]);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(ho, dout, sizeof(ho), cudaMemcpyDeviceToHost);
        double maxerr = 0, maxrow = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 256; ++n) {
                maxerr = fmax(maxerr, fabs((double)ho[m * 256 + n] - hb[n]));
                maxrow = fmax(maxrow, fabs((double)ho[m * 256 + n] - ho[n]));
            }
        printf("A(lbo=%d,sbo=%d) B(lbo=%d,sbo=%d): max |D - bias| = %.3e, max row-to-row difference = %.3e %s\n", c[0], c[1], c[2], c[3],
               maxerr, maxrow, e == cudaSuccess ? "" : cudaGetErrorString(e));
        for (int m : {0, 5, 8, 64, 127}) {
            printf("   D[%3d][0,1,7,8,9,100,255] =", m);
            for (int n : {0, 1, 7, 8, 9, 100, 255}) printf(" %8.4f", ho[m * 256 + n]);
            printf("\n");
        }
        printf("   bias      [0,1,7,8,9,100,255] =");
        for (int n : {0, 1, 7, 8, 9, 100, 255}) printf(" %8.4f", hb[n]);
        printf("\n");
    }
    return 0;
}
