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
// This file holds two kernels: sample_fine_kernel (any Nc multiple of 32, Nf a power of two) as described above
// (~1.9 k warp instructions per ray: the sort network makes it issue-bound), and sample_fine_fast_kernel for the
// shapes the reference runs at, which replaces step 3 by a checked linear placement (~600-780 warp instructions
// per ray; see the comment above it). HBM-bound by design (1.5-2 KB per ray at 64/128), issue-bound in practice.
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
                   const float* __restrict__ t_coarse, const float* __restrict__ u_fine, uint64_t seed,
                   const int64_t* __restrict__ step_dev, int64_t ray0,
                   float* __restrict__ t_sorted, int32_t* __restrict__ piece_idxs, float* __restrict__ cdf_out,
                   float* __restrict__ t_fine_out) {
    if (step_dev) seed ^= (uint64_t)__ldg(step_dev + 1);      // device-resident step counter (CUDA-graph replays)
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


// ---------------------------------------------------------------------------------------------
// Fast path for the shapes the reference is run at (N_coarse a power of two: 64/128 and 128/256).
// Same pdf/cdf/search/inversion arithmetic as above; what changes is how the ascending concat is
// produced. The register bitonic network + two rank-merge searches (≈1.2 k warp instructions per ray)
// are replaced by a placement that is linear in the number of samples:
//   * fine samples: bucket by floor(u*Nf) (u is i.i.d. uniform, so bucket occupancy is Poisson(1)
//     whatever the pdf looks like), smem-atomic arrival slots, one exclusive scan of the Nf counters,
//     rank inside the (tiny) bucket by direct comparison of t;
//   * number of coarse samples below a fine sample: its bin index plus one compare (each bin holds
//     exactly one coarse sample);
//   * coarse samples: the slots the fine samples left empty, in order.
// SORTED variant (no uniforms given): the kernel draws the order statistics of Nf uniforms directly, u arrives
// ascending across (lane, e) and the rank among the fine samples is simply the index - no buckets, no atomics.
// t is monotone in u only up to fp32 rounding across bin boundaries, and t_coarse is caller data, so
// the placement is CHECKED (no empty slot left, output ascending) and a warp whose check fails redoes
// the ray with the generic full bitonic sort: the result is always the exact sort.
constexpr uint32_t kHoleBits = 0xFFFFFFFFu;   // a NaN pattern no sample can carry through the check

template <int NC, int NF, bool SORTED>
struct alignas(16) FastWarpSmem {
    static constexpr int S = NC + NF;
    static constexpr int P2 = (S <= 256) ? 256 : ((S <= 512) ? 512 : 1024);
    float out[P2];        // merged samples (the generic fallback sorts the padded concat here)
    float tc[NC];         // coarse samples
    float tmp[SORTED ? 4 : NF];        // fine samples in bucket order       (explicit-u variant only)
    int cnt[SORTED ? 4 : NF];          // u-bucket occupancy
    int base[SORTED ? 4 : NF];         // (exclusive prefix of cnt) << 16 | cnt
    float pdf[NC];
    float edges[NC + 4];
    // cdf[k] lives at word k + (k >> 5): entries k and k+32 fall into different banks, which makes every step
    // of the lock-step binary search conflict-free (step j only ever reads entries that differ by multiples of 2^(j+1))
    float cdf[NC + NC / 32 + 4];
};
__host__ __device__ constexpr int cdf_slot(int k) { return k + (k >> 5); }

template <int NC, int NF, bool SORTED>
__global__ void __launch_bounds__(kSamplerWarps * 32)
sample_fine_fast_kernel(int64_t B, const float* __restrict__ bin_weights, const float* __restrict__ bin_edges,
                        const float* __restrict__ t_coarse, const float* __restrict__ u_fine, uint64_t seed,
                        const int64_t* __restrict__ step_dev, int64_t ray0, float* __restrict__ t_sorted,
                        int32_t* __restrict__ piece_idxs, float* __restrict__ cdf_out, float* __restrict__ t_fine_out) {
    if (step_dev) seed ^= (uint64_t)__ldg(step_dev + 1);      // device-resident step counter (CUDA-graph replays)
    static_assert((NC & (NC - 1)) == 0 && NC >= 32, "fast path: N_coarse must be a power of two");
    static_assert((NF & (NF - 1)) == 0 && NF >= 128, "fast path: N_fine must be a power of two >= 128");
    constexpr int EC = NC / 32, EF = NF / 32, S = NC + NF, ES = S / 32;
    static_assert(S % 32 == 0 && ES % 2 == 0, "fast path: 32 | (Nc+Nf)");
    using Smem = FastWarpSmem<NC, NF, SORTED>;
    constexpr int P2 = Smem::P2;
    __shared__ Smem sm_all[kSamplerWarps];
    constexpr unsigned FULL = 0xffffffffu;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ray = (int64_t)blockIdx.x * kSamplerWarps + warp;
    if (ray >= B) return;   // whole warp exits together; no block-level barriers below
    Smem& sm = sm_all[warp];

    // ---- loads: everything this ray needs from HBM is requested before the first use
    float w[EC], tcv[EC], u[EF];
    load_row<EC>(w, bin_weights + ray * NC + lane * EC);
    load_row<EC>(tcv, t_coarse + ray * NC + lane * EC);
    {
        const float* ep = bin_edges + ray * (NC + 1);    // rows of Nc+1 floats: not vector-aligned
        float ev[EC];
#pragma unroll
        for (int j = 0; j < EC; ++j) ev[j] = __ldg(ep + lane + 32 * j);
        if (lane == 0) sm.edges[NC] = __ldg(ep + NC);
#pragma unroll
        for (int j = 0; j < EC; ++j) sm.edges[lane + 32 * j] = ev[j];
    }
    if constexpr (!SORTED) {
        load_row<EF>(u, u_fine + ray * NF + lane * EF);
    } else {
        // No uniforms given: the reference draws NF i.i.d. U[0,1) per ray and sorts the samples they produce
        // (ray_utils.py:355,385). Their order statistics are generated directly instead - normalised partial sums of
        // NF+1 i.i.d. exponentials (Renyi's representation: same joint distribution) - so u arrives ascending across
        // (lane, e) and nothing has to be sorted. Philox counters as everywhere: (global ray id, block of 4, stream 1).
        float c[EF], extra = 0.f;
#pragma unroll
        for (int e0 = 0; e0 < EF; e0 += 4) {
            const uint64_t gr = (uint64_t)(ray0 + ray);
            const uint4 r = philox4x32_10(make_uint4((uint32_t)gr, (uint32_t)(gr >> 32), (uint32_t)((lane * EF + e0) >> 2), 1u),
                                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
            const uint32_t wd[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)   // v in (0,1] from the top 24 bits; -log2 v >= 0 (the log base cancels below)
                c[e0 + q] = -__log2f((float)((wd[q] >> 8) + 1u) * 5.9604644775390625e-8f);
            if (e0 == 0)                  // the (NF+1)-th exponential from the bits not used above (lane 31's is taken)
                extra = -__log2f((float)(((r.x & 0xffu) | ((r.y & 0xffu) << 8) | ((r.z & 0xffu) << 16)) + 1u) * 5.9604644775390625e-8f);
        }
#pragma unroll
        for (int e = 1; e < EF; ++e) c[e] += c[e - 1];
        float incl = c[EF - 1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        float excl = __shfl_up_sync(FULL, incl, 1);
        if (lane == 0) excl = 0.f;
        const float total = fmaxf(__shfl_sync(FULL, incl, 31) + __shfl_sync(FULL, extra, 31), 1e-30f);
        const float inv = __frcp_rn(total);
#pragma unroll
        for (int e = 0; e < EF; ++e) {
            // partial sum; the lane's last one IS the scan value the next lane starts from, the others are capped by
            // it, so the sequence is non-decreasing whatever the rounding of the tree scan
            const float part = (e == EF - 1) ? incl : fminf(excl + c[e], incl);
            u[e] = fminf(part * inv, 0.99999994f);
        }
    }
    store_row<EC>(sm.tc + lane * EC, tcv);
    {
        if constexpr (!SORTED) {
            float z[EF];
#pragma unroll
            for (int e = 0; e < EF; ++e) z[e] = 0.f;       // int 0 == float +0 bit pattern
            store_row<EF>(reinterpret_cast<float*>(sm.cnt) + lane * EF, z);
        }
        float h[ES];
#pragma unroll
        for (int j = 0; j < ES; ++j) h[j] = __uint_as_float(kHoleBits);
        store_row<ES>(sm.out + lane * ES, h);
    }
    __syncwarp();

    // ---- pdf / cdf (utils/ray_utils.py:335-345); identical arithmetic and summation order to the generic kernel
    float pv[EC], width[EC];
    {
        float ed[EC + 1];
#pragma unroll
        for (int e = 0; e <= EC; ++e) ed[e] = sm.edges[lane * EC + e];
        float local = 0.f;
#pragma unroll
        for (int e = 0; e < EC; ++e) {
            w[e] = __fadd_rn(w[e], 1e-5f);
            width[e] = __fsub_rn(ed[e + 1], ed[e]);
            local = __fadd_rn(local, __fmul_rn(w[e], width[e]));
        }
        float denom = local;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) denom = __fadd_rn(denom, __shfl_xor_sync(FULL, denom, o));
        float run = 0.f, cl[EC];
#pragma unroll
        for (int e = 0; e < EC; ++e) {
            pv[e] = __fdiv_rn(w[e], denom);
            run = __fadd_rn(run, __fmul_rn(pv[e], width[e]));
            cl[e] = run;
        }
        float incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            float v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl = __fadd_rn(incl, v);
        }
        float excl = __shfl_up_sync(FULL, incl, 1);
        if (lane == 0) excl = 0.f;
#pragma unroll
        for (int e = 0; e < EC; ++e) {
            sm.cdf[cdf_slot(lane * EC + e + 1)] = __fadd_rn(excl, cl[e]);
            sm.pdf[lane * EC + e] = pv[e];
        }
        if (lane == 0) sm.cdf[0] = 0.f;
    }
    __syncwarp();
    if (cdf_out)
        for (int k = lane; k <= NC; k += 32) cdf_out[ray * (NC + 1) + k] = sm.cdf[cdf_slot(k)];

    // ---- searchsorted(side='right') over the NC-1 inner edges cdf[1..NC-1] + inversion (:363-376)
    int idx[EF];
    float tf[EF];
    {
        // number of inner edges <= u; NC-1 = 2^k - 1 edges, so no bound checks. Byte-address form (one LDS with an
        // immediate offset, one compare, two conditional moves per step), the EF searches of a lane interleaved.
        // The last edge that compared <= u IS cdf[idx] (cdf[0] = 0), so the cdf gather of the inversion is free.
        const uint32_t cdf0 = (uint32_t)__cvta_generic_to_shared(sm.cdf);
        uint32_t a[EF];
        float cleft[EF];
#pragma unroll
        for (int e = 0; e < EF; ++e) { a[e] = cdf0; cleft[e] = 0.f; }
#pragma unroll
        for (int step = NC / 2; step > 0; step >>= 1) {
            const uint32_t off = 4u * (uint32_t)cdf_slot(step);
#pragma unroll
            for (int e = 0; e < EF; ++e) {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a[e] + off) : "memory");
                const bool take = v <= u[e];
                a[e] += take ? off : 0u;
                cleft[e] = take ? v : cleft[e];
            }
        }
#pragma unroll
        for (int e = 0; e < EF; ++e) {
            const uint32_t wd = (a[e] - cdf0) >> 2;        // = idx + idx/32
            const int pos = (int)(wd - wd / 33u);
            idx[e] = pos;
            float p = sm.pdf[pos];
            const float mask = p < 1e-8f ? 0.f : 1.f;
            p = fmaxf(p, 1e-8f);
            tf[e] = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(u[e], cleft[e]), p), mask), sm.edges[pos]);
        }
    }
    if (piece_idxs) {
#pragma unroll
        for (int e = 0; e < EF; ++e) piece_idxs[ray * NF + lane * EF + e] = idx[e];
    }
    if (t_fine_out) store_row<EF>(t_fine_out + ray * NF + lane * EF, tf);

    // ---- is t_coarse ascending? (it is when it comes from the stratified sampler)
    bool ok = true;
    {
        const float nxt = __shfl_down_sync(FULL, tcv[0], 1);
#pragma unroll
        for (int e = 0; e + 1 < EC; ++e) ok = ok && (tcv[e] <= tcv[e + 1]);
        if (lane < 31) ok = ok && (tcv[EC - 1] <= nxt);
    }
    if (__all_sync(FULL, ok)) {
        int frank[EF];   // rank of the sample among the fine samples
        if constexpr (SORTED) {
            // u ascending => t ascending, up to rounding across bin boundaries (checked below)
#pragma unroll
            for (int e = 0; e < EF; ++e) frank[e] = lane * EF + e;
        } else {
            // ---- fine samples: rank = (#fine in lower u-buckets) + (rank by t inside the bucket)
            int key[EF], slot[EF];
    #pragma unroll
            for (int e = 0; e < EF; ++e) {
                int k = (int)(u[e] * (float)NF);
                k = min(max(k, 0), NF - 1);
                key[e] = k;
                slot[e] = atomicAdd(&sm.cnt[k], 1);
            }
            __syncwarp();
            int excl, M;
            {
                int c[EF], loc = 0, cmax = 0;
    #pragma unroll
                for (int e = 0; e < EF; ++e) c[e] = sm.cnt[lane * EF + e];
                int ex[EF];
    #pragma unroll
                for (int e = 0; e < EF; ++e) { ex[e] = loc; loc += c[e]; cmax = max(cmax, c[e]); }
                int incl = loc;
    #pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                excl = incl - loc;
    #pragma unroll
                for (int e = 0; e < EF; ++e) sm.base[lane * EF + e] = ((excl + ex[e]) << 16) | c[e];
                M = __reduce_max_sync(FULL, cmax);
            }
            __syncwarp();
            int b[EF], n[EF], r[EF];
    #pragma unroll
            for (int e = 0; e < EF; ++e) {
                const int bc = sm.base[key[e]];
                b[e] = bc >> 16;
                n[e] = bc & 0xffff;
                r[e] = 0;
                sm.tmp[b[e] + slot[e]] = tf[e];
            }
            __syncwarp();
            // rank inside the bucket: compare with the OTHER members only. The load is predicated (s-th member exists and
            // is not this sample), so lanes whose bucket is exhausted put no traffic on the shared-memory pipe; an
            // unloaded `o` stays NaN and compares false both ways.
            const uint32_t tmp0 = (uint32_t)__cvta_generic_to_shared(sm.tmp);
            for (int s = 0; s < M; ++s) {
    #pragma unroll
                for (int e = 0; e < EF; ++e) {
                    float o = __uint_as_float(0x7fc00000u);
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "setp.lt.s32 p, %2, %3;\n\t"
                        "setp.ne.and.s32 p, %2, %4, p;\n\t"
                        "@p ld.shared.f32 %0, [%1];\n\t}"
                        : "+f"(o)
                        : "r"(tmp0 + 4u * (uint32_t)(b[e] + s)), "r"(s), "r"(n[e]), "r"(slot[e])
                        : "memory");
                    r[e] += ((o < tf[e]) | ((o == tf[e]) & (s < slot[e]))) ? 1 : 0;
                }
            }
#pragma unroll
            for (int e = 0; e < EF; ++e) frank[e] = b[e] + r[e];
        }
        // ---- merged position = fine rank + #coarse below; bin idx holds exactly one coarse sample
#pragma unroll
        for (int e = 0; e < EF; ++e) {
            const int below = idx[e] + (sm.tc[idx[e]] < tf[e] ? 1 : 0);
            sm.out[frank[e] + below] = tf[e];
        }
        __syncwarp();
        // ---- coarse samples fill the empty slots in order; then verify
        float ov[ES];
        load_row<ES>(ov, sm.out + lane * ES);
        int h = 0;
#pragma unroll
        for (int j = 0; j < ES; ++j) h += (__float_as_uint(ov[j]) == kHoleBits) ? 1 : 0;
        int hin = h;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(FULL, hin, o);
            if (lane >= o) hin += v;
        }
        const int total = __shfl_sync(FULL, hin, 31);
        if (total == NC) {   // warp-uniform
            int hp = hin - h;      // < NC wherever a hole is left, because total == NC
            const uint32_t tc0 = (uint32_t)__cvta_generic_to_shared(sm.tc);
#pragma unroll
            for (int j = 0; j < ES; ++j) {     // predicated load: only the lanes that own a hole read
                const uint32_t hole = (__float_as_uint(ov[j]) == kHoleBits) ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "setp.ne.u32 p, %2, 0;\n\t"
                    "@p ld.shared.f32 %0, [%1];\n\t}"
                    : "+f"(ov[j])
                    : "r"(tc0 + 4u * (uint32_t)hp), "r"(hole)
                    : "memory");
                hp += (int)hole;
            }
            const float nxt = __shfl_down_sync(FULL, ov[0], 1);
            bool asc = true;
#pragma unroll
            for (int j = 0; j + 1 < ES; ++j) asc = asc && (ov[j] <= ov[j + 1]);
            if (lane < 31) asc = asc && (ov[ES - 1] <= nxt);
            if (__all_sync(FULL, asc)) {
                store_row<ES>(t_sorted + ray * S + lane * ES, ov);
                return;
            }
        }
    }
    // ---- generic path (t_coarse not ascending, or the placement check failed): bitonic sort of the
    // padded concat in shared memory
    __syncwarp();
    for (int k = lane; k < NC; k += 32) sm.out[k] = sm.tc[k];
#pragma unroll
    for (int e = 0; e < EF; ++e) sm.out[NC + lane * EF + e] = tf[e];
    for (int k = S + lane; k < P2; k += 32) sm.out[k] = __int_as_float(0x7f800000);
    __syncwarp();
    for (int k = 2; k <= P2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < P2; i += 32) {
                int p = i ^ j;
                if (p > i) {
                    float a = sm.out[i], c = sm.out[p];
                    bool up = (i & k) == 0;
                    if ((a > c) == up) { sm.out[i] = c; sm.out[p] = a; }
                }
            }
            __syncwarp();
        }
    }
    for (int k = lane; k < S; k += 32) t_sorted[ray * S + k] = sm.out[k];
}

}  // namespace nb

using namespace nb;

extern "C" {

int nerfb200_sample_fine(int64_t B, int Nc, int Nf, const float* bin_weights, const float* bin_edges,
                         const float* t_coarse, const float* u_fine, uint64_t seed, const int64_t* step_state, int64_t ray0,
                         float* t_sorted, int32_t* piece_idxs, float* cdf, float* t_fine, void* stream) {
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
    cudaStream_t st = (cudaStream_t)stream;
    // fast path: the reference's shapes, vector-aligned rows (cudaMalloc'd tensors always are)
    const bool aligned = (((uintptr_t)bin_weights | (uintptr_t)t_coarse | (uintptr_t)t_sorted | (uintptr_t)u_fine |
                           (uintptr_t)t_fine) & 15) == 0;
    if (aligned && ((Nc == 64 && Nf == 128) || (Nc == 128 && Nf == 256))) {
        unsigned g = (unsigned)((B + kSamplerWarps - 1) / kSamplerWarps);
#define NB_LAUNCH_FAST(NCv, NFv, SRT)                                                                                   \
    sample_fine_fast_kernel<NCv, NFv, SRT><<<g, kSamplerWarps * 32, 0, st>>>(B, bin_weights, bin_edges, t_coarse, u_fine, \
                                                                             seed, step_state, ray0, t_sorted, piece_idxs, cdf, t_fine)
        if (Nc == 64) {
            if (u_fine) NB_LAUNCH_FAST(64, 128, false); else NB_LAUNCH_FAST(64, 128, true);
        } else {
            if (u_fine) NB_LAUNCH_FAST(128, 256, false); else NB_LAUNCH_FAST(128, 256, true);
        }
#undef NB_LAUNCH_FAST
        NB_LAUNCH_CHECK();
        return 0;
    }
    int S = Nc + Nf, P2 = 1;
    while (P2 < S) P2 <<= 1;
    size_t smem = (size_t)kSamplerWarps * ((Nc + 1) * 2 + Nc * 2 + Nf + P2) * sizeof(float);
    unsigned grid = (unsigned)((B + kSamplerWarps - 1) / kSamplerWarps);
#define NB_LAUNCH_SF(EFv)                                                                                      \
    do {                                                                                                       \
        if (smem > 48 * 1024)                                                                                  \
            NB_CUDA(cudaFuncSetAttribute(sample_fine_kernel<EFv>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem));                                                          \
        sample_fine_kernel<EFv><<<grid, kSamplerWarps * 32, smem, st>>>(B, Nc, bin_weights, bin_edges, t_coarse, \
                                                                         u_fine, seed, step_state, ray0, t_sorted, piece_idxs, \
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
