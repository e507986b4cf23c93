// Hierarchical (inverse-transform) sampler: create_input_batch_fine_model
// (utils/ray_utils.py:276-406). One warp per ray.
//
//   1. pdf/cdf: lane l owns bins [l*EC, (l+1)*EC); denominator by shuffle tree, inclusive CDF by a
//      lane-serial sum + shuffle scan; cdf/pdf/edges staged in shared memory (per-warp slice).
//   2. searchsorted(side='right') over the Nc-1 inner edges: branch-free binary search in smem,
//      then the reference's inversion with the pdf<1e-8 mask (separate mul/add, no FMA).
//   3. sort(concat(t_coarse, t_fine)): the Nf fine samples are sorted by a register bitonic
//      network (EF per lane, shuffles across lanes); t_coarse is already ascending when it comes
//      from the stratified sampler, so the concat-sort is finished by a rank merge (two binary
//      searches) and a scatter through shared memory. If t_coarse is NOT ascending the warp
//      falls back to a full bitonic sort of the padded concat in shared memory (same result).
//
// HBM-bound by design (1800 B/ray at 64/128); in practice the sort network makes it issue-bound.
#include "common.cuh"

namespace nb {

constexpr int kSamplerWarps = 4;

// Branch-free binary searches over a shared-memory array a[0, n): every lane runs the same log2 steps.
// upper_bound: #{a[i] <= v};  lower_bound: #{a[i] < v}.  p2 = largest power of two <= n.
__device__ __forceinline__ int upper_bound_smem(const float* a, int n, int p2, float v) {
    int pos = 0;
    for (int step = p2; step > 0; step >>= 1) {
        const int nxt = pos + step;
        if (nxt <= n && a[nxt - 1] <= v) pos = nxt;
    }
    return pos;
}
__device__ __forceinline__ int lower_bound_smem(const float* a, int n, int p2, float v) {
    int pos = 0;
    for (int step = p2; step > 0; step >>= 1) {
        const int nxt = pos + step;
        if (nxt <= n && a[nxt - 1] < v) pos = nxt;
    }
    return pos;
}

__device__ __forceinline__ void cmpswap(float& a, float& b, bool asc) {
    float lo = fminf(a, b), hi = fmaxf(a, b);
    a = asc ? lo : hi;
    b = asc ? hi : lo;
}

// Bitonic sort of 32*EF values, element index i = lane*EF + e, ascending.
template <int EF>
__device__ __forceinline__ void warp_bitonic_sort(float (&v)[EF], int lane) {
    constexpr int P = 32 * EF;
#pragma unroll
    for (int k = 2; k <= P; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= EF) {
                const int lj = j / EF;  // partner lane distance
#pragma unroll
                for (int e = 0; e < EF; ++e) {
                    int i = lane * EF + e;
                    float o = __shfl_xor_sync(0xffffffffu, v[e], lj);
                    bool asc = (i & k) == 0;
                    bool lower = (i & j) == 0;
                    v[e] = (lower == asc) ? fminf(v[e], o) : fmaxf(v[e], o);
                }
            } else {
#pragma unroll
                for (int e = 0; e < EF; ++e) {
                    if ((e & j) == 0) {
                        int i = lane * EF + e;
                        cmpswap(v[e], v[e | j], (i & k) == 0);
                    }
                }
            }
        }
    }
}

template <int EF>
__global__ void __launch_bounds__(kSamplerWarps * 32)
sample_fine_kernel(int64_t B, int Nc, const float* __restrict__ bin_weights, const float* __restrict__ bin_edges,
                   const float* __restrict__ t_coarse, const float* __restrict__ u_fine, uint64_t seed, int64_t ray0,
                   float* __restrict__ t_sorted, int32_t* __restrict__ piece_idxs, float* __restrict__ cdf_out,
                   float* __restrict__ t_fine_out) {
    constexpr int Nf = 32 * EF;
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ray = (int64_t)blockIdx.x * kSamplerWarps + warp;
    if (ray >= B) return;   // whole warp exits together; no block-level barriers below
    const int S = Nc + Nf;
    int P2 = 1;
    while (P2 < S) P2 <<= 1;
    // per-warp slices
    const int per_warp = (Nc + 1) * 2 + Nc * 2 + Nf + P2;
    float* s_edges = smem + warp * per_warp;   // Nc+1
    float* s_cdf = s_edges + (Nc + 1);         // Nc+1
    float* s_pdf = s_cdf + (Nc + 1);           // Nc
    float* s_tc = s_pdf + Nc;                  // Nc
    float* s_tf = s_tc + Nc;                   // Nf (sorted fine)
    float* s_out = s_tf + Nf;                  // P2 (>= Nc+Nf)

    const int EC = Nc >> 5;  // bins per lane (Nc is a multiple of 32)
    for (int k = lane; k <= Nc; k += 32) s_edges[k] = __ldg(bin_edges + ray * (Nc + 1) + k);
    for (int k = lane; k < Nc; k += 32) s_tc[k] = __ldg(t_coarse + ray * Nc + k);
    __syncwarp();

    // ---- pdf / cdf (utils/ray_utils.py:335-345)
    float local = 0.f;
    for (int e = 0; e < EC; ++e) {
        int k = lane * EC + e;
        float w = __fadd_rn(__ldg(bin_weights + ray * Nc + k), 1e-5f);
        float width = __fsub_rn(s_edges[k + 1], s_edges[k]);
        s_pdf[k] = w;                       // temporarily w'
        local = __fadd_rn(local, __fmul_rn(w, width));
    }
    float denom = local;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) denom = __fadd_rn(denom, __shfl_xor_sync(0xffffffffu, denom, o));
    float run = 0.f;
    for (int e = 0; e < EC; ++e) {
        int k = lane * EC + e;
        float p = __fdiv_rn(s_pdf[k], denom);
        float width = __fsub_rn(s_edges[k + 1], s_edges[k]);
        s_pdf[k] = p;
        run = __fadd_rn(run, __fmul_rn(p, width));
        s_cdf[k + 1] = run;                 // lane-local inclusive, fixed up below
    }
    float incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl = __fadd_rn(incl, v);
    }
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 0.f;
    for (int e = 0; e < EC; ++e) {
        int k = lane * EC + e;
        s_cdf[k + 1] = __fadd_rn(excl, s_cdf[k + 1]);
    }
    if (lane == 0) s_cdf[0] = 0.f;
    __syncwarp();
    if (cdf_out)
        for (int k = lane; k <= Nc; k += 32) cdf_out[ray * (Nc + 1) + k] = s_cdf[k];

    // ---- uniforms
    float u[EF];
    if (u_fine) {
#pragma unroll
        for (int e = 0; e < EF; ++e) u[e] = __ldg(u_fine + ray * Nf + lane * EF + e);
    } else {
#pragma unroll
        for (int e0 = 0; e0 < EF; e0 += 4) {
            float4 r = philox_uniform4(seed, (uint64_t)(ray0 + ray), (uint32_t)((lane * EF + e0) >> 2), 1u);
            float rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (e0 + q < EF) u[e0 + q] = rr[(EF >= 4) ? q : ((lane * EF + e0 + q) & 3)];
        }
    }

    // ---- searchsorted(side='right') over cdf[1..Nc-1] + inversion (utils/ray_utils.py:363-376)
    int p2c = 1;
    while (p2c * 2 <= Nc) p2c <<= 1;            // largest power of two <= Nc
    const int p2e = (p2c == Nc) ? (p2c >> 1) : p2c;   // ... and <= Nc - 1 (the inner edges)
    float tf[EF];
#pragma unroll
    for (int e = 0; e < EF; ++e) {
        const int idx = upper_bound_smem(s_cdf + 1, Nc - 1, p2e, u[e]);   // number of inner edges <= u, in [0, Nc-1]
        float p = s_pdf[idx];
        float mask = p < 1e-8f ? 0.f : 1.f;
        p = fmaxf(p, 1e-8f);
        float tv = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(u[e], s_cdf[idx]), p), mask), s_edges[idx]);
        tf[e] = tv;
        if (piece_idxs) piece_idxs[ray * Nf + lane * EF + e] = idx;
        if (t_fine_out) t_fine_out[ray * Nf + lane * EF + e] = tv;
    }

    // ---- is t_coarse ascending? (it is when it comes from the stratified sampler)
    bool ok = true;
    for (int k = lane; k + 1 < Nc; k += 32) ok = ok && (s_tc[k] <= s_tc[k + 1]);
    const bool coarse_sorted = __all_sync(0xffffffffu, ok);

    if (coarse_sorted) {
        warp_bitonic_sort<EF>(tf, lane);
#pragma unroll
        for (int e = 0; e < EF; ++e) s_tf[lane * EF + e] = tf[e];
        __syncwarp();
        // fine element with rank r goes to r + #{coarse < value}
#pragma unroll
        for (int e = 0; e < EF; ++e) s_out[lane * EF + e + lower_bound_smem(s_tc, Nc, p2c, tf[e])] = tf[e];
        // coarse element k goes to k + #{fine <= value}   (Nf = 32 EF is a power of two)
        for (int k = lane; k < Nc; k += 32) {
            const float v = s_tc[k];
            s_out[k + upper_bound_smem(s_tf, Nf, Nf, v)] = v;
        }
    } else {
        // generic path: bitonic sort of the padded concat in shared memory
        for (int k = lane; k < Nc; k += 32) s_out[k] = s_tc[k];
#pragma unroll
        for (int e = 0; e < EF; ++e) s_out[Nc + lane * EF + e] = tf[e];
        for (int k = S + lane; k < P2; k += 32) s_out[k] = __int_as_float(0x7f800000);
        __syncwarp();
        for (int k = 2; k <= P2; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = lane; i < P2; i += 32) {
                    int p = i ^ j;
                    if (p > i) {
                        float a = s_out[i], b = s_out[p];
                        bool asc = (i & k) == 0;
                        if ((a > b) == asc) { s_out[i] = b; s_out[p] = a; }
                    }
                }
                __syncwarp();
            }
        }
    }
    __syncwarp();
    for (int k = lane; k < S; k += 32) t_sorted[ray * S + k] = s_out[k];
}

}  // namespace nb

using namespace nb;

extern "C" {

int nerfb200_sample_fine(int64_t B, int Nc, int Nf, const float* bin_weights, const float* bin_edges,
                         const float* t_coarse, const float* u_fine, uint64_t seed, int64_t ray0, float* t_sorted,
                         int32_t* piece_idxs, float* cdf, float* t_fine, void* stream) {
    NB_CHECK_ARG(B >= 0, "sample_fine: negative ray count");
    if (!(Nc >= 32 && Nc <= 256 && Nc % 32 == 0)) {
        set_error("sample_fine: N_coarse must be a multiple of 32 in [32,256], got %d", Nc);
        return NERFB200_ENOTSUP;
    }
    if (!(Nf == 32 || Nf == 64 || Nf == 128 || Nf == 256 || Nf == 512)) {
        set_error("sample_fine: N_fine must be one of 32,64,128,256,512, got %d", Nf);
        return NERFB200_ENOTSUP;
    }
    if (B == 0) return 0;
    NB_CHECK_ARG(bin_weights && bin_edges && t_coarse && t_sorted, "sample_fine: NULL pointer");
    int S = Nc + Nf, P2 = 1;
    while (P2 < S) P2 <<= 1;
    size_t smem = (size_t)kSamplerWarps * ((Nc + 1) * 2 + Nc * 2 + Nf + P2) * sizeof(float);
    unsigned grid = (unsigned)((B + kSamplerWarps - 1) / kSamplerWarps);
    cudaStream_t st = (cudaStream_t)stream;
#define NB_LAUNCH_SF(EFv)                                                                                      \
    do {                                                                                                       \
        if (smem > 48 * 1024)                                                                                  \
            NB_CUDA(cudaFuncSetAttribute(sample_fine_kernel<EFv>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem));                                                          \
        sample_fine_kernel<EFv><<<grid, kSamplerWarps * 32, smem, st>>>(B, Nc, bin_weights, bin_edges, t_coarse, \
                                                                         u_fine, seed, ray0, t_sorted, piece_idxs, \
                                                                         cdf, t_fine);                        \
    } while (0)
    switch (Nf / 32) {
        case 1: NB_LAUNCH_SF(1); break;
        case 2: NB_LAUNCH_SF(2); break;
        case 4: NB_LAUNCH_SF(4); break;
        case 8: NB_LAUNCH_SF(8); break;
        default: NB_LAUNCH_SF(16); break;
    }
    NB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
