// Shared helpers for the nerfb200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <atomic>

#include "../../include/nerfb200.h"

namespace nb {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;   // kernels launched by this library (process-wide)

#define NB_CHECK_ARG(cond, ...)                        \
    do {                                               \
        if (!(cond)) {                                 \
            nb::set_error(__VA_ARGS__);                \
            return NERFB200_EINVAL;                    \
        }                                              \
    } while (0)

#define NB_CUDA(expr)                                                                  \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            nb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                          __FILE__, __LINE__);                                         \
            return (int)_e;                                                            \
        }                                                                              \
    } while (0)

#define NB_LAUNCH_CHECK()                 \
    do {                                  \
        nb::g_launches.fetch_add(1);      \
        NB_CUDA(cudaGetLastError());      \
    } while (0)

constexpr int kNumVars = NERFB200_NUM_VARS_PER_MODEL;
constexpr int kParamsPerModel = NERFB200_PARAMS_PER_MODEL;

// Layer table in the order of Keras' `model.layers` / `trainable_variables` for the functional model of
// core/model.py:334-394: Keras sorts layers by decreasing depth from the outputs [rgb, sigma], ties by the
// output-first traversal index, which puts the two heads last, rgb before sigma (it is NOT creation order).
enum Layer { L0 = 0, L1, L2, L3, L4, L5, L6, L7, L8, L9, LRGB, LSIGMA, kNumLayers };

struct LayerDim { int fan_in, fan_out; };
__host__ __device__ constexpr LayerDim layer_dim(int l) {
    return l == L0 ? LayerDim{63, 256}
         : l == L5 ? LayerDim{319, 256}
         : l == LSIGMA ? LayerDim{256, 1}
         : l == L9 ? LayerDim{283, 128}
         : l == LRGB ? LayerDim{128, 3}
         : LayerDim{256, 256};
}
__host__ __device__ constexpr int kernel_offset(int l) {
    int off = 0;
    for (int i = 0; i < l; ++i) off += layer_dim(i).fan_in * layer_dim(i).fan_out + layer_dim(i).fan_out;
    return off;
}
__host__ __device__ constexpr int bias_offset(int l) {
    return kernel_offset(l) + layer_dim(l).fan_in * layer_dim(l).fan_out;
}
static_assert(kernel_offset(kNumLayers) == kParamsPerModel, "parameter count mismatch");

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG: counter = (ray id lo, ray id hi, block-of-4 index, stream id),
// key = seed. Keyed by the GLOBAL ray id so results do not depend on how rays are sharded
// across GPUs or chunks (SURVEY.md section 7 "RNG under sharding").
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
// uint32 -> fp32 in [0,1) with 23 random mantissa bits (same construction as TF's Philox uniform).
__device__ __forceinline__ float u32_to_unit_float(uint32_t x) {
    return __uint_as_float((x >> 9) | 0x3F800000u) - 1.0f;
}
__device__ __forceinline__ float4 philox_uniform4(uint64_t seed, uint64_t ray, uint32_t block, uint32_t stream_id) {
    uint4 r = philox4x32_10(make_uint4((uint32_t)ray, (uint32_t)(ray >> 32), block, stream_id),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    return make_float4(u32_to_unit_float(r.x), u32_to_unit_float(r.y), u32_to_unit_float(r.z),
                       u32_to_unit_float(r.w));
}

// ---------------------------------------------------------------------------------------------
// Row helpers: N consecutive floats moved as the widest vectors dividing N (the address must be
// aligned to that vector).
template <int N> struct VecOf;
template <> struct VecOf<1> { using type = float; };
template <> struct VecOf<2> { using type = float2; };
template <> struct VecOf<4> { using type = float4; };
// widest vector (in floats) dividing n
__host__ __device__ constexpr int vec_width(int n) { return n % 4 == 0 ? 4 : (n % 2 == 0 ? 2 : 1); }

// N consecutive floats, N*4-byte aligned groups of vec_width(N)
template <int N>
__device__ __forceinline__ void load_row(float (&dst)[N], const float* __restrict__ src) {
    constexpr int V = vec_width(N);
    using T = typename VecOf<V>::type;
#pragma unroll
    for (int i = 0; i < N / V; ++i) {
        T v = *reinterpret_cast<const T*>(src + i * V);
        const float* f = reinterpret_cast<const float*>(&v);
#pragma unroll
        for (int q = 0; q < V; ++q) dst[i * V + q] = f[q];
    }
}
template <int N>
__device__ __forceinline__ void store_row(float* __restrict__ dst, const float (&src)[N]) {
    constexpr int V = vec_width(N);
    using T = typename VecOf<V>::type;
#pragma unroll
    for (int i = 0; i < N / V; ++i) {
        T v;
        float* f = reinterpret_cast<float*>(&v);
#pragma unroll
        for (int q = 0; q < V; ++q) f[q] = src[i * V + q];
        *reinterpret_cast<T*>(dst + i * V) = v;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// Keras OptimizerV2 Adam._resource_apply_dense (non-amsgrad; core/model.py:413-418), shared by
// adam_kernel (optim.cu) and the fused exchange + Adam kernel (peer.cu).
// lr_t = lr(iterations) * sqrt(1 - b2^t) / (1 - b1^t), t = iterations + 1; lr = ExponentialDecay(5e-4, 500000, 0.1), staircase=False
__host__ __device__ inline float adam_lr_t(int64_t iterations) {
    const double beta1 = 0.9, beta2 = 0.999;
    const double t = (double)(iterations + 1);
    const float lr = (float)(5e-4 * pow(0.1, (double)iterations / 500000.0));
    return lr * (float)sqrt(1.0 - pow(beta2, t)) / (float)(1.0 - pow(beta1, t));
}
// m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr_t m / (sqrt(v) + eps): separate IEEE operations, no contraction
__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, float lr_t) {
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-7f;
    m = __fadd_rn(__fmul_rn(m, b1), __fmul_rn(g, 1.f - b1));
    v = __fadd_rn(__fmul_rn(v, b2), __fmul_rn(__fmul_rn(g, g), 1.f - b2));
    p = __fsub_rn(p, __fdiv_rn(__fmul_rn(lr_t, m), __fadd_rn(__fsqrt_rn(v), eps)));
}

}  // namespace nb
