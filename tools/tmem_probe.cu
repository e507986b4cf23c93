// Developer microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM for a few shapes / warp counts.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int X> __device__ __forceinline__ void ld(uint32_t a, uint32_t& acc);
template <> __device__ __forceinline__ void ld<32>(uint32_t a, uint32_t& acc) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
          "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]) : "r"(a));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= r[i];
}
template <> __device__ __forceinline__ void ld<64>(uint32_t a, uint32_t& acc) {
    uint32_t r[64];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
        "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
        : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
          "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]),
          "=r"(r[32]),"=r"(r[33]),"=r"(r[34]),"=r"(r[35]),"=r"(r[36]),"=r"(r[37]),"=r"(r[38]),"=r"(r[39]),"=r"(r[40]),"=r"(r[41]),"=r"(r[42]),"=r"(r[43]),"=r"(r[44]),"=r"(r[45]),"=r"(r[46]),"=r"(r[47]),
          "=r"(r[48]),"=r"(r[49]),"=r"(r[50]),"=r"(r[51]),"=r"(r[52]),"=r"(r[53]),"=r"(r[54]),"=r"(r[55]),"=r"(r[56]),"=r"(r[57]),"=r"(r[58]),"=r"(r[59]),"=r"(r[60]),"=r"(r[61]),"=r"(r[62]),"=r"(r[63]) : "r"(a));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 64; ++i) acc ^= r[i];
}
template <int X>
__global__ void __launch_bounds__(512, 1) probe(int warps, int iters, long long* out, uint32_t* sink) {
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_ptr;
    uint32_t acc = 0;
    long long t0 = clock64();
    if (warp < warps) {
        const uint32_t base = tm + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128 % 512);
        for (int i = 0; i < iters; ++i) ld<X>(base + (uint32_t)((i * X) % 128), acc);
    }
    long long t1 = clock64();
    if (threadIdx.x % 32 == 0 && warp < warps) out[blockIdx.x * 16 + warp] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}
int main() {
    long long* out; cudaMalloc(&out, 148 * 16 * sizeof(long long)); uint32_t* sink; cudaMalloc(&sink, 4);
    const int iters = 4096;
    long long h[16];
    for (int warps : {1, 4, 8, 16}) {
        probe<32><<<148, 512>>>(warps, iters, out, sink); cudaDeviceSynchronize();
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("x32 warps=%2d: %.1f cycles per ld per warp -> %.1f B/cycle/SM\n", warps, (double)mx / iters, (double)warps * iters * 32 * 32 * 4 / mx);
        probe<64><<<148, 512>>>(warps, iters, out, sink); cudaDeviceSynchronize();
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("x64 warps=%2d: %.1f cycles per ld per warp -> %.1f B/cycle/SM  (%s)\n", warps, (double)mx / iters, (double)warps * iters * 64 * 32 * 4 / mx, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
