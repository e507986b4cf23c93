// Volume-rendering integrator: sigma_to_alpha / compute_weights / post_process_model_output
// (utils/ray_utils.py:408-551), forward and backward.
//
// One warp per ray. Lane l owns the E = ceil(S/32) consecutive samples [l*E, (l+1)*E), so a warp's
// loads cover one contiguous S*4-byte (sigma, t) or S*12-byte (rgb) span of HBM. The exclusive
// transmittance product is a lane-local serial product followed by a 5-step shuffle scan of the
// 32 lane totals. HBM-bound: 24*S+20 bytes per ray.
#include "common.cuh"

namespace nb {

constexpr int kWarpsPerBlock = 8;

// Warp-cooperative copy of n floats global -> shared with cp.async (16-byte pieces when `vec`, i.e. both
// addresses 16-byte aligned and 4 | n; 4-byte pieces otherwise). MAXN >= n is the compile-time bound the
// vector loop is unrolled to. Completion: cp_async_wait_all + __syncwarp.
template <int MAXN>
__device__ __forceinline__ void warp_cp_async(float* sdst, const float* __restrict__ gsrc, int n, bool vec, int lane) {
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(sdst);
    if (vec) {
        const int n4 = n >> 2;
#pragma unroll
        for (int j = 0; j < (MAXN / 4 + 31) / 32; ++j) {
            const int i = lane + 32 * j;
            if (i < n4)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + 16u * i), "l"(gsrc + 4 * i) : "memory");
        }
    } else {
#pragma unroll 1
        for (int i = lane; i < n; i += 32)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s0 + 4u * i), "l"(gsrc + i) : "memory");
    }
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// per-warp staging buffer, in floats: t (+1 closing value), sigma (reused for the weights), rgb
__host__ __device__ constexpr int composite_warp_floats(int E) { return (32 * E + 4) + 32 * E + 96 * E; }

// Forward. The ray's t, sigma and rgb spans are staged in shared memory with fully coalesced 16-byte
// cp.async copies (every 32-byte sector of HBM is requested exactly once, all requests of the ray in
// flight together); the lanes then read their E consecutive samples from shared memory with vector
// loads (conflict-free for the strides that occur: 2, 6, 12 and 3x those). The weights go back through
// the sigma slots and leave as coalesced 16-byte stores.
// FULL: the warp's span is exactly 32*E samples (64, 128, 192, 256, 384 ... samples per ray with one ray per warp), so
// every bound is a compile-time constant. RPW (rays per warp, 1 or 2; 2 needs FULL): with RPW = 2 a warp takes two
// ADJACENT rays of 16*E samples each - lanes 0-15 the first, 16-31 the second. The two rays are one contiguous span
// of HBM, so staging is unchanged; the scan and the reductions run over 16-lane segments (one shuffle step fewer) and
// the per-warp overhead (setup, copies, scan, five reductions, output) is paid once per two rays. Used for S = 64,
// where that overhead, not HBM, bounded the one-ray-per-warp form (241 warp instructions per ray, 69 % of HBM peak).
template <int E, bool FULL, int RPW = 1>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_fwd_kernel(int64_t B, int S_arg, const float* __restrict__ sigma, const float* __restrict__ rgb,
                     const float* __restrict__ t_vals, int white_bg, float* __restrict__ weights,
                     float* __restrict__ pred_rgb, float* __restrict__ pred_depth, float* __restrict__ acc_map) {
    static_assert(RPW == 1 || (RPW == 2 && FULL), "two rays per warp only with compile-time sizes");
    constexpr int W = 32 / RPW;                       // lanes per ray
    extern __shared__ __align__(16) float comp_smem[];
    const int lane = threadIdx.x & 31;
    const int64_t ray0 = ((int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5)) * RPW;   // first ray of this warp
    if (ray0 >= B) return;   // whole warp exits together; no block-level barriers below
    float* s_t = comp_smem + (threadIdx.x >> 5) * composite_warp_floats(E);
    float* s_sg = s_t + (32 * E + 4);
    float* s_rgb = s_sg + 32 * E;
    const int S = FULL ? W * E : S_arg;               // samples per ray
    const int nrays = (RPW == 2 && ray0 + 1 < B) ? 2 : 1;
    const int span = nrays * S;                       // samples this warp stages
    const int64_t base = ray0 * S;
    const int s0 = lane * E;                          // position in the warp's span
    const int ls = (lane % W) * E;                    // position in the lane's ray
    const int64_t ray = ray0 + lane / W;
    const bool live = (lane / W) < nrays;             // false for the second half of an odd last warp

    const bool vec = ((S & 3) == 0) &&
                     ((((uintptr_t)sigma | (uintptr_t)rgb | (uintptr_t)t_vals | (uintptr_t)weights) & 15) == 0);
    warp_cp_async<32 * E>(s_t, t_vals + base, span, vec, lane);
    warp_cp_async<32 * E>(s_sg, sigma + base, span, vec, lane);
    warp_cp_async<96 * E>(s_rgb, rgb + 3 * base, 3 * span, vec, lane);
    cp_async_wait_all();
    __syncwarp();

    float t[E + 1], sg[E];
    {
        float tt[E];
        load_row<E>(tt, s_t + s0);
        load_row<E>(sg, s_sg + s0);
#pragma unroll
        for (int e = 0; e < E; ++e) t[e] = tt[e];
        t[E] = s_t[s0 + E];   // first t of the next lane closes this lane's last interval (unused past S-1)
    }

    float alpha[E], f[E];
    float lane_prod = 1.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int s = ls + e;
        const bool in = live && s < S;
        float delta = (s == S - 1) ? 1e10f : __fsub_rn(t[e + 1], t[e]);           // :462-468
        float a = __fsub_rn(1.f, expf(-__fmul_rn(sg[e], delta)));                  // :423
        a = in ? a : 0.f;
        alpha[e] = a;
        f[e] = in ? __fadd_rn(__fsub_rn(1.f, a), 1e-10f) : 1.f;                   // :480
        lane_prod *= f[e];
    }
    // exclusive scan (product) of lane totals, per ray
    float incl = lane_prod;
#pragma unroll
    for (int o = 1; o < W; o <<= 1) {
        float v = __shfl_up_sync(0xffffffffu, incl, o, W);
        if ((lane % W) >= o) incl *= v;
    }
    float T = __shfl_up_sync(0xffffffffu, incl, 1, W);
    if ((lane % W) == 0) T = 1.f;

    float cr = 0.f, cg = 0.f, cb = 0.f, dep = 0.f, acc = 0.f;
    float c[3 * E], wv[E];
    load_row<3 * E>(c, s_rgb + 3 * s0);
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int s = ls + e;
        float w = alpha[e] * T;
        T *= f[e];
        wv[e] = w;
        if (live && s < S) {
            cr += w * c[3 * e + 0];
            cg += w * c[3 * e + 1];
            cb += w * c[3 * e + 2];
            dep += w * t[e];
            acc += w;
        }
    }
    if (weights) store_row<E>(s_sg + s0, wv);   // this lane's own slots: already consumed above
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) {       // butterfly over the ray's lanes
        cr += __shfl_xor_sync(0xffffffffu, cr, o);
        cg += __shfl_xor_sync(0xffffffffu, cg, o);
        cb += __shfl_xor_sync(0xffffffffu, cb, o);
        dep += __shfl_xor_sync(0xffffffffu, dep, o);
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
    }
    if (weights) {
        __syncwarp();
        float* wout = weights + base;
        if (vec) {
#pragma unroll
            for (int j = 0; j < (8 * E + 31) / 32; ++j) {
                const int i = lane + 32 * j;
                if (i < (span >> 2)) reinterpret_cast<float4*>(wout)[i] = reinterpret_cast<const float4*>(s_sg)[i];
            }
        } else {
#pragma unroll 1
            for (int i = lane; i < span; i += 32) wout[i] = s_sg[i];
        }
    }
    if ((lane % W) == 0 && live) {
        if (white_bg) {                                                            // :542-544
            float bg = __fsub_rn(1.f, acc);
            cr += bg; cg += bg; cb += bg;
        }
        pred_rgb[3 * ray + 0] = cr;
        pred_rgb[3 * ray + 1] = cg;
        pred_rgb[3 * ray + 2] = cb;
        pred_depth[ray] = dep;
        acc_map[ray] = acc;
    }
}

// Backward w.r.t. sigma and per-sample rgb. With g_i = sum_ch dC_ch*(c_i,ch - bg):
//   dL/dalpha_i = g_i*T_i - (sum_{k>i} g_k*w_k) / f_i      (TF's cumprod gradient: reverse cumsum / x)
//   dL/dsigma_i = dL/dalpha_i * delta_i * exp(-sigma_i*delta_i)
//   dL/dc_i,ch  = w_i * dC_ch
template <int E>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_bwd_kernel(int64_t B, int S, const float* __restrict__ sigma, const float* __restrict__ rgb,
                     const float* __restrict__ t_vals, int white_bg, const float* __restrict__ d_pred,
                     float* __restrict__ d_sigma, float* __restrict__ d_rgb) {
    const int lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (ray >= B) return;
    const int64_t base = ray * S;
    const int s0 = lane * E;
    const float dr = __ldg(d_pred + 3 * ray + 0), dg = __ldg(d_pred + 3 * ray + 1), db = __ldg(d_pred + 3 * ray + 2);
    const float bg = white_bg ? 1.f : 0.f;

    float t[E + 1], sg[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        int s = s0 + e;
        t[e] = s < S ? __ldg(t_vals + base + s) : 0.f;
        sg[e] = s < S ? __ldg(sigma + base + s) : 0.f;
    }
    t[E] = __shfl_down_sync(0xffffffffu, t[0], 1);

    float alpha[E], f[E], ex[E], delta[E];
    float lane_prod = 1.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        int s = s0 + e;
        delta[e] = (s == S - 1) ? 1e10f : __fsub_rn(t[e + 1], t[e]);
        ex[e] = s < S ? expf(-__fmul_rn(sg[e], delta[e])) : 1.f;
        alpha[e] = s < S ? __fsub_rn(1.f, ex[e]) : 0.f;
        f[e] = s < S ? __fadd_rn(__fsub_rn(1.f, alpha[e]), 1e-10f) : 1.f;
        lane_prod *= f[e];
    }
    float incl = lane_prod;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl *= v;
    }
    float T = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) T = 1.f;

    float Tn[E], gw[E], g[E];
    float lane_gw = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        int s = s0 + e;
        Tn[e] = T;
        float w = alpha[e] * T;
        T *= f[e];
        float gi = 0.f;
        if (s < S) {
            const float* c = rgb + 3 * (base + s);
            gi = dr * (__ldg(c + 0) - bg) + dg * (__ldg(c + 1) - bg) + db * (__ldg(c + 2) - bg);
            float* o = d_rgb + 3 * (base + s);
            o[0] = w * dr; o[1] = w * dg; o[2] = w * db;
        }
        g[e] = gi;
        gw[e] = gi * w;
        lane_gw += gw[e];
    }
    // exclusive suffix sum of lane totals (reverse scan)
    float sincl = lane_gw;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float v = __shfl_down_sync(0xffffffffu, sincl, o);
        if (lane + o < 32) sincl += v;
    }
    float suffix = __shfl_down_sync(0xffffffffu, sincl, 1);
    if (lane == 31) suffix = 0.f;
#pragma unroll
    for (int e = E - 1; e >= 0; --e) {
        int s = s0 + e;
        float dalpha = g[e] * Tn[e] - suffix / f[e];
        suffix += gw[e];
        if (s < S) d_sigma[base + s] = dalpha * delta[e] * ex[e];
    }
}

// Training form of the integrator: forward, loss and backward of a ray in ONE launch (NeRF.train_step,
// core/model.py:148-170: post_process_model_output -> MeanSquaredError -> tape.gradient back to sigma and rgb).
// = composite_fwd_kernel<E, FULL, RPW> (same lane mapping, same operations in the same order: the forward outputs are
// bit-identical to the render path's, which keeps the hierarchical samples of a training forward identical to a render
// forward's), the arithmetic of mse_loss_grad_kernel on the ray's three channels, and composite_bwd_kernel on values that
// are still in registers (alpha, f, T, exp, the colours): three launches and a second pass over t / sigma / rgb become
// one. At 512 rays per GPU (one rank's share of a data-parallel step on 8 GPUs) the three kernels were 3-6 us each,
// almost all of it launch and drain.
template <int E, bool FULL, int RPW = 1>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_train_kernel(int64_t B, int S_arg, const float* __restrict__ sigma, const float* __restrict__ rgb,
                       const float* __restrict__ t_vals, int white_bg, const float* __restrict__ rgb_gt, float inv_n,
                       float* __restrict__ weights, float* __restrict__ pred_rgb, float* __restrict__ pred_depth,
                       float* __restrict__ acc_map, float* __restrict__ d_sigma, float* __restrict__ d_rgb,
                       float* __restrict__ loss, float* __restrict__ metric) {
    static_assert(RPW == 1 || (RPW == 2 && FULL), "two rays per warp only with compile-time sizes");
    constexpr int W = 32 / RPW;                       // lanes per ray
    extern __shared__ __align__(16) float comp_smem[];
    const int lane = threadIdx.x & 31;
    const int64_t ray0 = ((int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5)) * RPW;   // first ray of this warp
    if (ray0 >= B) return;   // whole warp exits together; no block-level barriers below
    float* s_t = comp_smem + (threadIdx.x >> 5) * composite_warp_floats(E);
    float* s_sg = s_t + (32 * E + 4);
    float* s_rgb = s_sg + 32 * E;
    const int S = FULL ? W * E : S_arg;               // samples per ray
    const int nrays = (RPW == 2 && ray0 + 1 < B) ? 2 : 1;
    const int span = nrays * S;                       // samples this warp stages
    const int64_t base = ray0 * S;
    const int s0 = lane * E;                          // position in the warp's span
    const int ls = (lane % W) * E;                    // position in the lane's ray
    const int64_t ray = ray0 + lane / W;
    const bool live = (lane / W) < nrays;             // false for the second half of an odd last warp

    const bool vec = ((S & 3) == 0) &&
                     ((((uintptr_t)sigma | (uintptr_t)rgb | (uintptr_t)t_vals | (uintptr_t)weights) & 15) == 0);
    warp_cp_async<32 * E>(s_t, t_vals + base, span, vec, lane);
    warp_cp_async<32 * E>(s_sg, sigma + base, span, vec, lane);
    warp_cp_async<96 * E>(s_rgb, rgb + 3 * base, 3 * span, vec, lane);
    cp_async_wait_all();
    __syncwarp();

    // ---- forward (composite_fwd_kernel, operation for operation)
    float t[E + 1], sg[E];
    {
        float tt[E];
        load_row<E>(tt, s_t + s0);
        load_row<E>(sg, s_sg + s0);
#pragma unroll
        for (int e = 0; e < E; ++e) t[e] = tt[e];
        t[E] = s_t[s0 + E];   // first t of the next lane closes this lane's last interval (unused past S-1)
    }
    float alpha[E], f[E], ex[E], delta[E];
    float lane_prod = 1.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int s = ls + e;
        const bool in = live && s < S;
        delta[e] = (s == S - 1) ? 1e10f : __fsub_rn(t[e + 1], t[e]);               // :462-468
        ex[e] = expf(-__fmul_rn(sg[e], delta[e]));
        float a = __fsub_rn(1.f, ex[e]);                                           // :423
        a = in ? a : 0.f;
        alpha[e] = a;
        f[e] = in ? __fadd_rn(__fsub_rn(1.f, a), 1e-10f) : 1.f;                   // :480
        lane_prod *= f[e];
    }
    float incl = lane_prod;
#pragma unroll
    for (int o = 1; o < W; o <<= 1) {
        float v = __shfl_up_sync(0xffffffffu, incl, o, W);
        if ((lane % W) >= o) incl *= v;
    }
    float T = __shfl_up_sync(0xffffffffu, incl, 1, W);
    if ((lane % W) == 0) T = 1.f;

    float cr = 0.f, cg = 0.f, cb = 0.f, dep = 0.f, acc = 0.f;
    float c[3 * E], wv[E], Tn[E];
    load_row<3 * E>(c, s_rgb + 3 * s0);
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int s = ls + e;
        Tn[e] = T;
        float w = alpha[e] * T;
        T *= f[e];
        wv[e] = w;
        if (live && s < S) {
            cr += w * c[3 * e + 0];
            cg += w * c[3 * e + 1];
            cb += w * c[3 * e + 2];
            dep += w * t[e];
            acc += w;
        }
    }
    if (weights) store_row<E>(s_sg + s0, wv);   // this lane's own slots: already consumed above
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) {       // butterfly over the ray's lanes: every lane ends up with the ray's totals
        cr += __shfl_xor_sync(0xffffffffu, cr, o);
        cg += __shfl_xor_sync(0xffffffffu, cg, o);
        cb += __shfl_xor_sync(0xffffffffu, cb, o);
        dep += __shfl_xor_sync(0xffffffffu, dep, o);
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
    }
    if (weights) {
        __syncwarp();
        float* wout = weights + base;
        if (vec) {
#pragma unroll
            for (int j = 0; j < (8 * E + 31) / 32; ++j) {
                const int i = lane + 32 * j;
                if (i < (span >> 2)) reinterpret_cast<float4*>(wout)[i] = reinterpret_cast<const float4*>(s_sg)[i];
            }
        } else {
#pragma unroll 1
            for (int i = lane; i < span; i += 32) wout[i] = s_sg[i];
        }
    }
    if (white_bg) {                                                                // :542-544
        const float bg1 = __fsub_rn(1.f, acc);
        cr += bg1; cg += bg1; cb += bg1;
    }
    // ---- loss (mse_loss_grad_kernel's arithmetic on this ray's three channels; core/model.py:157-168)
    float dr = 0.f, dg = 0.f, db = 0.f;
    if (live) {
        const float e0 = cr - __ldg(rgb_gt + 3 * ray + 0), e1 = cg - __ldg(rgb_gt + 3 * ray + 1), e2 = cb - __ldg(rgb_gt + 3 * ray + 2);
        dr = 2.f * e0 * inv_n; dg = 2.f * e1 * inv_n; db = 2.f * e2 * inv_n;
        if ((lane % W) == 0) {
            pred_rgb[3 * ray + 0] = cr;
            pred_rgb[3 * ray + 1] = cg;
            pred_rgb[3 * ray + 2] = cb;
            pred_depth[ray] = dep;
            acc_map[ray] = acc;
            const float sq = e0 * e0 + e1 * e1 + e2 * e2;
            atomicAdd(loss, sq * inv_n);
            if (metric) {                                                          // PSNRMetric.update_state, core/ops.py:204-220
                atomicAdd(metric, sq);
                if (ray == 0) atomicAdd(metric + 1, (float)B);
            }
        }
    }
    // ---- backward (composite_bwd_kernel's arithmetic, on the values above; the scans run over the ray's W lanes)
    const float bg = white_bg ? 1.f : 0.f;
    float gw[E], g[E];
    float lane_gw = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int s = ls + e;
        float gi = 0.f;
        if (live && s < S) {
            gi = dr * (c[3 * e + 0] - bg) + dg * (c[3 * e + 1] - bg) + db * (c[3 * e + 2] - bg);
            float* o = d_rgb + 3 * (ray * S + s);
            o[0] = wv[e] * dr; o[1] = wv[e] * dg; o[2] = wv[e] * db;
        }
        g[e] = gi;
        gw[e] = gi * wv[e];
        lane_gw += gw[e];
    }
    float sincl = lane_gw;                      // inclusive suffix sum of the lane totals within the ray's lanes
#pragma unroll
    for (int o = 1; o < W; o <<= 1) {
        float v = __shfl_down_sync(0xffffffffu, sincl, o, W);
        if ((lane % W) + o < W) sincl += v;
    }
    float suffix = __shfl_down_sync(0xffffffffu, sincl, 1, W);
    if ((lane % W) == W - 1) suffix = 0.f;
#pragma unroll
    for (int e = E - 1; e >= 0; --e) {
        const int s = ls + e;
        const float dalpha = g[e] * Tn[e] - suffix / f[e];
        suffix += gw[e];
        if (live && s < S) d_sigma[ray * S + s] = dalpha * delta[e] * ex[e];
    }
}

static inline int pick_E(int S) { return (S + 31) / 32; }

#define NB_DISPATCH_E(Eval, ...)                                   \
    switch (Eval) {                                                \
        case 1: { constexpr int E = 1; __VA_ARGS__; } break;       \
        case 2: { constexpr int E = 2; __VA_ARGS__; } break;       \
        case 3: { constexpr int E = 3; __VA_ARGS__; } break;       \
        case 4: { constexpr int E = 4; __VA_ARGS__; } break;       \
        case 5: case 6: { constexpr int E = 6; __VA_ARGS__; } break;   \
        case 7: case 8: { constexpr int E = 8; __VA_ARGS__; } break;   \
        case 9: case 10: case 11: case 12: { constexpr int E = 12; __VA_ARGS__; } break; \
        case 13: case 14: case 15: case 16: { constexpr int E = 16; __VA_ARGS__; } break; \
        default: { constexpr int E = 32; __VA_ARGS__; } break;     \
    }

}  // namespace nb

using namespace nb;

extern "C" {

int nerfb200_composite_fwd(int64_t B, int S, const float* sigma, const float* rgb, const float* t_vals, int white_bg,
                           float* weights, float* pred_rgb, float* pred_depth, float* acc_map, void* stream) {
    NB_CHECK_ARG(B >= 0 && S >= 2 && S <= 1024, "composite_fwd: need 2 <= S <= 1024, got S=%d", S);
    if (B == 0) return 0;
    NB_CHECK_ARG(sigma && rgb && t_vals && pred_rgb && pred_depth && acc_map, "composite_fwd: NULL pointer");
    if (S == 64) {
        // two adjacent rays per warp, a half-warp each (E = 4 samples per lane)
        constexpr int smem2 = kWarpsPerBlock * composite_warp_floats(4) * (int)sizeof(float);
        const int64_t warps = (B + 1) / 2;
        composite_fwd_kernel<4, true, 2><<<(unsigned)((warps + kWarpsPerBlock - 1) / kWarpsPerBlock), kWarpsPerBlock * 32, smem2,
                                           (cudaStream_t)stream>>>(B, S, sigma, rgb, t_vals, white_bg, weights, pred_rgb, pred_depth,
                                                                   acc_map);
        NB_LAUNCH_CHECK();
        return 0;
    }
    unsigned grid = (unsigned)((B + kWarpsPerBlock - 1) / kWarpsPerBlock);
    NB_DISPATCH_E(pick_E(S), {
        constexpr int smem = kWarpsPerBlock * composite_warp_floats(E) * (int)sizeof(float);
        if (S == 32 * E) {
            if (smem > 48 * 1024) NB_CUDA(cudaFuncSetAttribute(composite_fwd_kernel<E, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            composite_fwd_kernel<E, true><<<grid, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
                B, S, sigma, rgb, t_vals, white_bg, weights, pred_rgb, pred_depth, acc_map);
        } else {
            if (smem > 48 * 1024) NB_CUDA(cudaFuncSetAttribute(composite_fwd_kernel<E, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            composite_fwd_kernel<E, false><<<grid, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
                B, S, sigma, rgb, t_vals, white_bg, weights, pred_rgb, pred_depth, acc_map);
        }
    });
    NB_LAUNCH_CHECK();
    return 0;
}

int nerfb200_composite_bwd(int64_t B, int S, const float* sigma, const float* rgb, const float* t_vals, int white_bg,
                           const float* d_pred_rgb, float* d_sigma, float* d_rgb, void* stream) {
    NB_CHECK_ARG(B >= 0 && S >= 2 && S <= 1024, "composite_bwd: need 2 <= S <= 1024, got S=%d", S);
    if (B == 0) return 0;
    NB_CHECK_ARG(sigma && rgb && t_vals && d_pred_rgb && d_sigma && d_rgb, "composite_bwd: NULL pointer");
    unsigned grid = (unsigned)((B + kWarpsPerBlock - 1) / kWarpsPerBlock);
    NB_DISPATCH_E(pick_E(S), (composite_bwd_kernel<E><<<grid, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
                                 B, S, sigma, rgb, t_vals, white_bg, d_pred_rgb, d_sigma, d_rgb)));
    NB_LAUNCH_CHECK();
    return 0;
}

int nerfb200_composite_train(int64_t B, int S, const float* sigma, const float* rgb, const float* t_vals, int white_bg,
                             const float* rgb_gt, int64_t B_global, float* weights, float* pred_rgb, float* pred_depth,
                             float* acc_map, float* d_sigma, float* d_rgb, float* loss, float* metric, void* stream) {
    NB_CHECK_ARG(B >= 0 && S >= 2 && S <= 1024, "composite_train: need 2 <= S <= 1024, got S=%d", S);
    NB_CHECK_ARG(B_global >= B, "composite_train: B_global < B");
    if (B == 0) return 0;
    NB_CHECK_ARG(sigma && rgb && t_vals && rgb_gt && pred_rgb && pred_depth && acc_map && d_sigma && d_rgb && loss,
                 "composite_train: NULL pointer");
    const float inv_n = 1.0f / (float)(B_global * 3);
    if (S == 64) {
        // the lane mapping of composite_fwd for 64 samples: two adjacent rays per warp, a half-warp each
        constexpr int smem2 = kWarpsPerBlock * composite_warp_floats(4) * (int)sizeof(float);
        const int64_t warps = (B + 1) / 2;
        composite_train_kernel<4, true, 2><<<(unsigned)((warps + kWarpsPerBlock - 1) / kWarpsPerBlock), kWarpsPerBlock * 32, smem2,
                                             (cudaStream_t)stream>>>(B, S, sigma, rgb, t_vals, white_bg, rgb_gt, inv_n, weights,
                                                                     pred_rgb, pred_depth, acc_map, d_sigma, d_rgb, loss, metric);
        NB_LAUNCH_CHECK();
        return 0;
    }
    unsigned grid = (unsigned)((B + kWarpsPerBlock - 1) / kWarpsPerBlock);
    NB_DISPATCH_E(pick_E(S), {
        constexpr int smem = kWarpsPerBlock * composite_warp_floats(E) * (int)sizeof(float);
        if (S == 32 * E) {
            if (smem > 48 * 1024) NB_CUDA(cudaFuncSetAttribute(composite_train_kernel<E, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            composite_train_kernel<E, true><<<grid, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
                B, S, sigma, rgb, t_vals, white_bg, rgb_gt, inv_n, weights, pred_rgb, pred_depth, acc_map, d_sigma, d_rgb, loss, metric);
        } else {
            if (smem > 48 * 1024) NB_CUDA(cudaFuncSetAttribute(composite_train_kernel<E, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            composite_train_kernel<E, false><<<grid, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
                B, S, sigma, rgb, t_vals, white_bg, rgb_gt, inv_n, weights, pred_rgb, pred_depth, acc_map, d_sigma, d_rgb, loss, metric);
        }
    });
    NB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
