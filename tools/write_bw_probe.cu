// Pure-write HBM bandwidth by store flavour (how close can a write-only kernel get to the copy bandwidth?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/write_bw_probe tools/write_bw_probe.cu
// Variants: cudaMemsetAsync; st.global.v4 (default / .cs streaming / .wt); cp.async.bulk shared->global in 16 KB pieces
// (the store the training kernels use for their stashes), with and without an L2 evict_first hint; and a mixed
// kernel that reads one buffer while writing another (what a copy does), for reference.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>
__global__ void stg_kernel(uint4* __restrict__ dst, size_t n16) {
    const uint4 v = make_uint4(1, 2, 3, threadIdx.x);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        if (MODE == 0) dst[i] = v;
        else if (MODE == 1) __stcs(dst + i, v);
        else __stwt(dst + i, v);
    }
}

__global__ void copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = __ldcs(src + i);
}

// one CTA = one SM; every iteration ships PIECE bytes of shared memory to global with ONE bulk store per stage
template <bool HINT>
__global__ void __launch_bounds__(128, 1) bulk_kernel(uint8_t* __restrict__ dst, size_t npieces, int piece) {
    extern __shared__ __align__(128) uint8_t sm[];
    for (int i = threadIdx.x; i < piece / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t pol = 0;
        if (HINT) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
        int inflight = 0;
        for (size_t p = blockIdx.x; p < npieces; p += gridDim.x) {
            uint8_t* g = dst + p * (size_t)piece;
            if (HINT)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(g), "r"(s), "r"(piece), "l"(pol) : "memory");
            else
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(s), "r"(piece) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++inflight >= 8) { asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory"); }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// The training kernels' store pattern: per SM two "slots"; each slot ships its 64 KB activation buffer with one bulk
// store, waits until the engine has READ the buffer (wait_group.read 0), then spends `idle` cycles refilling it (the
// epilogue) before it can ship again. Achieved bandwidth vs idle time tells how much of the stash cost is the
// store/refill serialisation and how much a deeper staging buffer could recover.
__global__ void __launch_bounds__(64, 1) bursty_kernel(uint8_t* __restrict__ dst, size_t npieces, int idle) {
    extern __shared__ __align__(128) uint8_t sm[];
    constexpr int kBuf = 65536;
    const int slot = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 2 * kBuf / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm + slot * kBuf);
        for (size_t p = (size_t)blockIdx.x * 2 + slot; p < npieces; p += (size_t)gridDim.x * 2) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + p * (size_t)kBuf), "r"(s), "r"(kBuf) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            const long long t0 = clock64();
            while (clock64() - t0 < idle) { }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main() {
    const size_t bytes = (size_t)4 << 30;
    uint8_t *a, *b;
    CK(cudaMalloc(&a, bytes));
    CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(a, 1, bytes));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t n16 = bytes / 16;
    auto report = [&](const char* name, float ms, double payload) { printf("%-58s %8.3f ms  %7.1f GB/s\n", name, ms, payload / ms / 1e6); };
#define TIME(name, payload, ...)                                   \
    do {                                                           \
        for (int w = 0; w < 2; ++w) { __VA_ARGS__; }               \
        CK(cudaDeviceSynchronize());                               \
        cudaEventRecord(e0);                                       \
        for (int w = 0; w < 5; ++w) { __VA_ARGS__; }               \
        cudaEventRecord(e1);                                       \
        CK(cudaDeviceSynchronize());                               \
        float ms; cudaEventElapsedTime(&ms, e0, e1);               \
        report(name, ms / 5, (double)(payload));                   \
    } while (0)

    TIME("cudaMemsetAsync", bytes, cudaMemsetAsync(b, 7, bytes));
    for (int mult : {4, 8, 16}) {
        char nm[96];
        snprintf(nm, sizeof nm, "st.global.v4 default, grid %d x SMs, 512 thr", mult);
        TIME(nm, bytes, (stg_kernel<0><<<sms * mult, 512>>>((uint4*)b, n16)));
    }
    TIME("st.global.cs.v4 (streaming), grid 8 x SMs", bytes, (stg_kernel<1><<<sms * 8, 512>>>((uint4*)b, n16)));
    TIME("st.global.wt.v4 (write-through), grid 8 x SMs", bytes, (stg_kernel<2><<<sms * 8, 512>>>((uint4*)b, n16)));
    for (int piece : {4096, 16384, 65536}) {
        char nm[96];
        cudaFuncSetAttribute(bulk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        cudaFuncSetAttribute(bulk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        snprintf(nm, sizeof nm, "cp.async.bulk smem->global, %d B pieces, 1 CTA/SM", piece);
        TIME(nm, bytes, (bulk_kernel<false><<<sms, 128, piece>>>(b, bytes / piece, piece)));
        snprintf(nm, sizeof nm, "cp.async.bulk + L2 evict_first hint, %d B pieces", piece);
        TIME(nm, bytes, (bulk_kernel<true><<<sms, 128, piece>>>(b, bytes / piece, piece)));
    }
    cudaFuncSetAttribute(bursty_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536);
    for (int idle : {0, 1000, 2000, 3000, 5000}) {
        char nm[96];
        snprintf(nm, sizeof nm, "2 slots/SM: 64 KB store, wait read, refill %d cycles", idle);
        TIME(nm, bytes, (bursty_kernel<<<sms, 64, 2 * 65536>>>(b, bytes / 65536, idle)));
    }
    TIME("copy (ld.cs + st), payload counted once", bytes, (copy_kernel<<<sms * 8, 512>>>((const uint4*)a, (uint4*)b, n16)));
    printf("(copy moves 2x its payload: read + write)\n");
    return 0;
}
