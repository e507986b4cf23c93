// Backward of the 8x256 MLP (tf.GradientTape over core/model.py:148-170) on the tensor cores.
//
// Three kernels per model and step, all tcgen05.mma with fp32 accumulation in TMEM and 16-bit operands:
//
//  1. bwd_data_pair_kernel -- per 128-row tile, the chain dZ_9 -> dZ_8 -> ... -> dZ_0 stays on chip like the
//     forward: dX = dZ . W^T (W^T streamed through the bulk-TMA ring), the epilogue applies the ReLU
//     bitmask stashed by the training forward and writes dZ_{l-1} as the next A operand. The rgb and
//     sigma heads (N = 3 and 1) are fp32 CUDA-core work in the prologue/epilogue. Every dZ_l leaves as
//     a chunk image (tc_layout.cuh) for kernel 2. No dX for the network inputs: sample positions are
//     constants for autodiff (stop_gradient, utils/ray_utils.py:377).
//  2. dw_kernel        -- dW_l = X_l^T . dZ_l, a reduction over ALL rows. One launch runs the 14 jobs of a
//     model side by side (job j owns a K-split of CTAs proportional to the bytes it streams); a CTA walks
//     its share of the tiles, bulk-loads the X_l and dZ_l chunk images and feeds them to the tensor core as
//     MN-major operands (K = rows); the [K_in x N_out] fp32 accumulator lives in TMEM for the whole pass
//     and is flushed once per CTA. Bias gradients (column sums of dZ_l) are accumulated by the otherwise
//     idle warps from the same shared-memory tiles. HBM bound (95 % of the measured peak).
//  3. reduce_grads_kernel -- one launch folds the per-CTA dumps of every job, in a fixed order, into the
//     flat fp32 gradient buffer (Keras kernel layout [in,out]).
#include "common.cuh"
#include "mlp.cuh"
#include "tc_common.cuh"
#include "tc_layout.cuh"

#include <stdlib.h>

namespace nb {

// =============================================================================================
// W^T image for backward-data: chunks [128 in-features x 64 out-features], K-major (K = out features).
// Jobs: B1 = dense_9 (bott part, K = 128), B2 = dense_8, B3..B9 = dense_7..dense_1 (dense_5: h4 part).
constexpr int kBwdJobs = 9;
struct BwdChunk { uint32_t gofs; uint8_t layer, nh, kc, pad; };
constexpr int kBwdMaxChunks = 72;
struct BwdTable { BwdChunk c[kBwdMaxChunks]; int n; int job_begin[kBwdJobs + 1]; uint32_t bytes; };

static BwdTable build_bwd_table() {
    BwdTable t{};
    const int layers[kBwdJobs] = {L9, L8, L7, L6, L5, L4, L3, L2, L1};
    uint32_t ofs = 0;
    int n = 0;
    for (int j = 0; j < kBwdJobs; ++j) {
        t.job_begin[j] = n;
        const int KC = j == 0 ? 2 : 4;
        for (int nh = 0; nh < 2; ++nh)
            for (int kc = 0; kc < KC; ++kc) {
                t.c[n++] = BwdChunk{ofs, (uint8_t)layers[j], (uint8_t)nh, (uint8_t)kc, 0};
                ofs += kChunkBytes;
            }
    }
    t.job_begin[kBwdJobs] = n;
    t.n = n;
    t.bytes = ofs;
    return t;
}
static const BwdTable& bwd_table() { static BwdTable t = build_bwd_table(); return t; }

__constant__ BwdChunk c_bwd_chunks[kBwdMaxChunks];
__constant__ int c_bwd_job_begin[kBwdJobs + 1];

template <typename T> __device__ __forceinline__ T cvt16(float v);
template <> __device__ __forceinline__ __nv_bfloat16 cvt16<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half cvt16<__half>(float v) { return __float2half_rn(v); }

template <typename T>
__global__ void pack_bwd_weights_kernel(const float* __restrict__ flat_params, uint8_t* __restrict__ img0, uint8_t* __restrict__ img1,
                                        int nchunks) {
    const float* P = flat_params + (int64_t)blockIdx.y * kParamsPerModel;      // blockIdx.y = model (coarse, fine)
    uint8_t* img = blockIdx.y ? img1 : img0;
    int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= (int64_t)nchunks * (kChunkBytes / 16)) return;
    const int ci = (int)(u / (kChunkBytes / 16));
    const BwdChunk ch = c_bwd_chunks[ci];
    const uint32_t local = (uint32_t)(u % (kChunkBytes / 16)) * 16;
    const int row = local >> 7, unit = ((local >> 4) & 7) ^ (row & 7);
    const LayerDim dim = layer_dim(ch.layer);
    const float* W = P + kernel_offset(ch.layer);
    const int n = ch.nh * 128 + row;           // in-feature (row of the Keras kernel)
    T vals[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = ch.kc * 64 + unit * 8 + e;   // out-feature (column of the Keras kernel)
        float v = (n < dim.fan_in && k < dim.fan_out) ? W[(int64_t)n * dim.fan_out + k] : 0.f;
        vals[e] = cvt16<T>(v);
    }
    *reinterpret_cast<uint4*>(img + ch.gofs + local) = *reinterpret_cast<uint4*>(vals);
}

// =============================================================================================
// 1. backward-data kernel
constexpr int kBStages = 4;
constexpr int kBSmemAct = 0;                                   // 2 x 64 KB: dZ of the current layer (A operand)
constexpr int kBSmemHead = kBSmemAct + 2 * 4 * kChunkBytes;    // 2 x 16 KB: head chunk (dZ_rgb, dZ_sigma)
constexpr int kBSmemRing = kBSmemHead + 2 * kChunkBytes;       // 4 x 16 KB: W^T ring
constexpr int kBSmemBar = kBSmemRing + kBStages * kChunkBytes;
constexpr int kBSmemWsig = kBSmemBar + 256;                    // 256 fp32: sigma kernel
constexpr int kBSmemTotal = kBSmemWsig + 1024;
static_assert(kBSmemTotal <= 232448, "exceeds the 227 KB dynamic shared memory limit");
constexpr int kBThreads = 320;
constexpr int kBProducerWarp = 8, kBMmaWarp = 9;

struct BwdParams {
    const uint8_t* wimg;
    const float* P;              // fp32 master parameters of this model (rgb / sigma heads)
    const uint8_t* stash;        // activation stash of the training forward
    uint8_t* gstash;             // out: gradient stash
    const float* d_rgb; const float* d_sigma;
    int64_t R; int num_tiles;
};

// One backward layer's epilogue for this thread's row: dZ_prev = (acc [+ dzs*wsig]) masked by the ReLU bitmask.
template <bool kHalf, bool kMask, bool kSigma>
__device__ __forceinline__ void bwd_epilogue_cols(uint32_t tmem_row, uint8_t* act, int row, const uint32_t (&mask)[8],
                                                  const float* s_wsig, float dzs, bool valid) {
    uint32_t r[2][32];
    tmem_ld32(tmem_row, r[0]);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        uint32_t (&rr)[32] = r[g & 1];
        tmem_ld_wait(rr);
        if (g + 1 < 8) tmem_ld32(tmem_row + (uint32_t)(32 * (g + 1)), r[(g + 1) & 1]);
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rr[i]);
        if (kSigma) {
            const float4* w4 = reinterpret_cast<const float4*>(s_wsig + 32 * g);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 w = w4[i];
                v[4 * i + 0] = fmaf(dzs, w.x, v[4 * i + 0]);
                v[4 * i + 1] = fmaf(dzs, w.y, v[4 * i + 1]);
                v[4 * i + 2] = fmaf(dzs, w.z, v[4 * i + 2]);
                v[4 * i + 3] = fmaf(dzs, w.w, v[4 * i + 3]);
            }
        }
        if (kMask) {
            const uint32_t m = valid ? mask[g] : 0u;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = (m >> mask_bit(i)) & 1u ? v[i] : 0.f;
        } else if (!valid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        uint8_t* chunk = act + (g >> 1) * kChunkBytes;
        const int u0 = (g & 1) * 4;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            uint4 o;
            o.x = pack2<kHalf>(v[8 * u + 0], v[8 * u + 1]);
            o.y = pack2<kHalf>(v[8 * u + 2], v[8 * u + 3]);
            o.z = pack2<kHalf>(v[8 * u + 4], v[8 * u + 5]);
            o.w = pack2<kHalf>(v[8 * u + 6], v[8 * u + 7]);
            *reinterpret_cast<uint4*>(chunk + swz(row, u0 + u)) = o;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 2-CTA variant of the backward-data kernel (tcgen05 cta_group::2, see mlp_tc_forward_pair_kernel): M = 256
// MMAs across the CTA pair, each CTA streams HALF of W^T, and every weight stage is used by both slots.
template <bool kHalf>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kBThreads, 1) bwd_data_pair_kernel(const BwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t sbase = smem_u32(smem);
    const uint32_t sbar = sbase + kBSmemBar;
    auto ring_full = [&](int s) { return sbar + 8 * s; };
    auto ring_empty = [&](int s) { return sbar + 8 * (kBStages + s); };
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kBSmemBar + 8 * (2 * kBStages + 4));
    float* s_wsig = reinterpret_cast<float*>(smem + kBSmemWsig);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kBStages; ++s) { mbar_init(ring_full(s), rank == 0 ? 2 : 1); mbar_init(ring_empty(s), 1); }
        for (int t = 0; t < 2; ++t) { mbar_init(sbar + 8 * (2 * kBStages + t), 2); mbar_init(sbar + 8 * (2 * kBStages + 2 + t), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 256) s_wsig[threadIdx.x] = p.P[kernel_offset(LSIGMA) + threadIdx.x];
    cluster_sync_all();
    if (warp == kBMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int num_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
    const int quads = (p.num_tiles + 3) >> 2;
    constexpr int fmt = kHalf ? 0 : 1;

    if (warp == kBProducerWarp) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int qd = cluster_id; qd < quads; qd += num_clusters)
                for (int j = 0; j < kBwdJobs; ++j) {
                    const int KC = j == 0 ? 2 : 4;
                    for (int kc = 0; kc < KC; ++kc) {          // loaded once, used by both slots
                        mbar_wait(ring_empty(stage), phase ^ 1);
                        mbar_expect_tx(ring_full(stage), kChunkBytes);
                        bulk_g2s(sbase + kBSmemRing + stage * kChunkBytes,
                                 p.wimg + c_bwd_chunks[c_bwd_job_begin[j] + (int)rank * KC + kc].gofs, kChunkBytes, ring_full(stage));
                        if (++stage == kBStages) { stage = 0; phase ^= 1; }
                    }
                }
        }
    } else if (warp == kBMmaWarp) {
        if (lane == 0 && rank == 1) {
            uint32_t stage = 0, phase = 0;
            for (int qd = cluster_id; qd < quads; qd += num_clusters)
                for (int j = 0; j < kBwdJobs; ++j) {
                    const int KC = j == 0 ? 2 : 4;
                    for (int kc = 0; kc < KC; ++kc) {
                        mbar_wait(ring_full(stage), phase);
                        mbar_arrive_cluster(mapa(ring_full(stage), 0));
                        if (++stage == kBStages) { stage = 0; phase ^= 1; }
                    }
                }
        } else if (rank == 0) {
            // converged warp, one elected lane issues (see mlp_tc_forward_pair_kernel)
            const uint32_t ring_lo = ((sbase + kBSmemRing) >> 4) & 0x3FFFu;
            constexpr uint32_t idesc = umma_idesc_pair(fmt, 256);
            uint32_t stage = 0, phase = 0, act_phase_bits = 0;
            for (int qd = cluster_id; qd < quads; qd += num_clusters) {
                const int nslots = (qd * 4 + 2 < p.num_tiles) ? 2 : 1;
#pragma unroll 1
                for (int j = 0; j < kBwdJobs; ++j) {
                    const int KC = j == 0 ? 2 : 4;
                    const uint32_t stage0 = stage, phase0 = phase;
#pragma unroll 1
                    for (int t = 0; t < nslots; ++t) {
                        mbar_wait_cluster(sbar + 8 * (2 * kBStages + t), (act_phase_bits >> t) & 1u);
                        act_phase_bits ^= 1u << t;
                        tc_fence_after();
                        const uint32_t act_lo = ((sbase + kBSmemAct + t * 4 * kChunkBytes) >> 4) & 0x3FFFu;
                        const uint32_t d = tmem_base + (uint32_t)(t * 256);
                        stage = stage0; phase = phase0;
                        const bool first_user = t == 0, last_user = t == nslots - 1;
                        // one election per (job, slot): the elected lane walks the K chunks alone (see the forward kernel)
                        if (elect_one_sync()) {
                            uint32_t st = stage, ph = phase;
#pragma unroll 1
                            for (int kc = 0; kc < KC; ++kc) {
                                if (first_user) { mbar_wait_cluster(ring_full(st), ph); tc_fence_after(); }
                                const uint32_t a_lo = act_lo + (uint32_t)(kc * 1024);
                                const uint32_t b_lo = ring_lo + st * (kChunkBytes >> 4);
                                umma_f16_pair(d, umma_desc_from_lo(a_lo), umma_desc_from_lo(b_lo), idesc, kc == 0 ? 0u : 1u);
                                umma_f16_pair(d, umma_desc_from_lo(a_lo + 2), umma_desc_from_lo(b_lo + 2), idesc, 1u);
                                umma_f16_pair(d, umma_desc_from_lo(a_lo + 4), umma_desc_from_lo(b_lo + 4), idesc, 1u);
                                umma_f16_pair(d, umma_desc_from_lo(a_lo + 6), umma_desc_from_lo(b_lo + 6), idesc, 1u);
                                if (last_user) umma_commit_pair(ring_empty(st));
                                if (kc == KC - 1) umma_commit_pair(sbar + 8 * (2 * kBStages + 2 + t));
                                if (++st == kBStages) { st = 0; ph ^= 1; }
                            }
                        }
                        __syncwarp();
                        phase ^= ((stage + (uint32_t)KC) / kBStages) & 1u;
                        stage = (stage + (uint32_t)KC) % kBStages;
                    }
                }
            }
        }
    } else if (warp < 8) {
        const int t = warp >> 2, q = warp & 3, row = q * 32 + lane;
        uint8_t* act = smem + kBSmemAct + t * 4 * kChunkBytes;
        uint8_t* head = smem + kBSmemHead + t * kChunkBytes;
        const uint32_t act_saddr = sbase + kBSmemAct + t * 4 * kChunkBytes, head_saddr = sbase + kBSmemHead + t * kChunkBytes;
        const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 256);
        const uint32_t act_ready_leader = mapa(sbar + 8 * (2 * kBStages + t), 0), acc_full = sbar + 8 * (2 * kBStages + 2 + t);
        uint32_t acc_phase = 0;
        uint8_t* pend_dst = nullptr; uint32_t pend_bytes = 0; bool pend_head = false; uint8_t* pend_head_dst = nullptr;

        for (int qd = cluster_id; qd < quads; qd += num_clusters) {
            if (qd * 4 + t * 2 >= p.num_tiles) continue;
            const int tile = qd * 4 + t * 2 + (int)rank;
            const bool real_tile = tile < p.num_tiles;                 // a dummy tile contributes zeros and stores nothing
            const uint8_t* tstash = p.stash + (size_t)(real_tile ? tile : 0) * kStashTileBytes;
            uint8_t* gst = p.gstash + (size_t)(real_tile ? tile : 0) * kGradTileBytes;
            const int64_t grow = (int64_t)tile * kTileRows + row;
            const bool valid = real_tile && grow < p.R;
            const float* outs = reinterpret_cast<const float*>(tstash + kStashOutOfs);
            float dzr[3] = {0.f, 0.f, 0.f}, dzs = 0.f;
            if (valid) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float y = outs[3 * row + c];
                    dzr[c] = __ldg(p.d_rgb + 3 * grow + c) * y * (1.f - y);
                }
                dzs = outs[3 * 128 + row] > 0.f ? __ldg(p.d_sigma + grow) : 0.f;
            }
            // previous tile's last stores must have read smem before it is overwritten
            named_bar_sync(1 + t, kTileRows);
            if (row == 0) {
                bool issued = false;
                if (pend_bytes) { bulk_s2g(pend_dst, act_saddr, pend_bytes); issued = true; }
                if (pend_head) { bulk_s2g(pend_head_dst, head_saddr, kChunkBytes); issued = true; }
                if (issued) { bulk_commit_group(); bulk_wait_read_all(); }
            }
            pend_bytes = 0; pend_head = false;
            named_bar_sync(1 + t, kTileRows);
            {
                uint4 o = make_uint4(pack2<kHalf>(dzr[0], dzr[1]), pack2<kHalf>(dzr[2], dzs), 0u, 0u);
                *reinterpret_cast<uint4*>(head + swz(row, 0)) = o;
#pragma unroll
                for (int u = 1; u < 8; ++u) *reinterpret_cast<uint4*>(head + swz(row, u)) = make_uint4(0u, 0u, 0u, 0u);
            }
            {
                const uint32_t* mrow = reinterpret_cast<const uint32_t*>(tstash + kStashMaskOfs) + (8 * 128 + row) * 8;
                const uint4 m4 = *reinterpret_cast<const uint4*>(mrow);
                const uint32_t mm[4] = {m4.x, m4.y, m4.z, m4.w};
                const float* Wrgb = p.P + kernel_offset(LRGB);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float v[32];
                    const uint32_t m = valid ? mm[g] : 0u;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int c = 32 * g + i;
                        float a = dzr[0] * __ldg(Wrgb + 3 * c) + dzr[1] * __ldg(Wrgb + 3 * c + 1) + dzr[2] * __ldg(Wrgb + 3 * c + 2);
                        v[i] = (m >> mask_bit(i)) & 1u ? a : 0.f;
                    }
                    uint8_t* chunk = act + (g >> 1) * kChunkBytes;
                    const int u0 = (g & 1) * 4;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        uint4 o;
                        o.x = pack2<kHalf>(v[8 * u + 0], v[8 * u + 1]);
                        o.y = pack2<kHalf>(v[8 * u + 2], v[8 * u + 3]);
                        o.z = pack2<kHalf>(v[8 * u + 4], v[8 * u + 5]);
                        o.w = pack2<kHalf>(v[8 * u + 6], v[8 * u + 7]);
                        *reinterpret_cast<uint4*>(chunk + swz(row, u0 + u)) = o;
                    }
                }
            }
            fence_proxy_async();
            if (real_tile) {
                pend_dst = gst + kGradChunkZ9 * kChunkBytes; pend_bytes = 2 * kChunkBytes;
                pend_head = true; pend_head_dst = gst + kGradChunkHead * kChunkBytes;
            }
            for (int j = 0; j < kBwdJobs; ++j) {
                named_bar_sync(1 + t, kTileRows);                 // this CTA's operand for job j is complete and fenced
                bool issued = false;
                if (row == 0) {
                    mbar_arrive_cluster(act_ready_leader);
                    if (pend_bytes) { bulk_s2g(pend_dst, act_saddr, pend_bytes); issued = true; }
                    if (pend_head) { bulk_s2g(pend_head_dst, head_saddr, kChunkBytes); issued = true; }
                    if (issued) bulk_commit_group();
                }
                pend_bytes = 0; pend_head = false;
                uint32_t mask[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                if (j >= 1) {
                    const uint4* mrow = reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(tstash + kStashMaskOfs) + ((8 - j) * 128 + row) * 8);
                    const uint4 a = mrow[0], b = mrow[1];
                    mask[0] = a.x; mask[1] = a.y; mask[2] = a.z; mask[3] = a.w;
                    mask[4] = b.x; mask[5] = b.y; mask[6] = b.z; mask[7] = b.w;
                }
                mbar_wait(acc_full, acc_phase);
                acc_phase ^= 1;
                tc_fence_after();
                if (row == 0 && issued) bulk_wait_read_all();
                named_bar_sync(1 + t, kTileRows);
                if (j == 0) bwd_epilogue_cols<kHalf, false, false>(tmem_row, act, row, mask, s_wsig, dzs, valid);
                else if (j == 1) bwd_epilogue_cols<kHalf, true, true>(tmem_row, act, row, mask, s_wsig, dzs, valid);
                else bwd_epilogue_cols<kHalf, true, false>(tmem_row, act, row, mask, s_wsig, dzs, valid);
                tc_fence_before();
                fence_proxy_async();
                if (real_tile) { pend_dst = gst + (j == 0 ? kGradChunkZ8 : grad_chunk_Z(8 - j)) * kChunkBytes; pend_bytes = 4 * kChunkBytes; }
            }
        }
        named_bar_sync(1 + t, kTileRows);
        if (row == 0) {
            if (pend_bytes) bulk_s2g(pend_dst, act_saddr, pend_bytes);
            if (pend_head) bulk_s2g(pend_head_dst, head_saddr, kChunkBytes);
            bulk_commit_group();
            bulk_wait_all();
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == kBMmaWarp) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

// =============================================================================================
// 2. weight-gradient kernel: dW = X^T . dZ with MN-major operands (K = rows)
struct DwJob {
    int x_chunk0, x_nchunks;      // X_l chunk images in the activation stash (1, 2 or 4 chunks of 64 features)
    int dz_chunk0, dz_nchunks;    // dZ_l chunk images in the gradient stash (1, 2 or 4 chunks of 64 columns)
    int layer;                    // Layer enum of the Keras kernel receiving the gradient
    int k_row0, k_rows;           // rows [k_row0, k_row0 + k_rows) of that kernel come from X features [0, k_rows)
    int n_col0, n_cols;           // kernel columns [0, n_cols) come from dZ columns [n_col0, n_col0 + n_cols)
    int bias_layer;               // Layer enum receiving colsum(dZ) (same column mapping), or -1
};
constexpr int kDwJobs = 14;
static void build_dw_jobs(DwJob* j) {
    int n = 0;
    j[n++] = DwJob{kStashChunkEncXyz, 1, grad_chunk_Z(0), 4, L0, 0, 63, 0, 256, L0};
    for (int l = 1; l <= 7; ++l) j[n++] = DwJob{stash_chunk_Y(l - 1), 4, grad_chunk_Z(l), 4, l, 0, 256, 0, 256, l};
    j[n++] = DwJob{kStashChunkEncXyz, 1, grad_chunk_Z(5), 4, L5, 256, 63, 0, 256, -1};
    j[n++] = DwJob{stash_chunk_Y(7), 4, kGradChunkZ8, 4, L8, 0, 256, 0, 256, L8};
    j[n++] = DwJob{stash_chunk_Y(7), 4, kGradChunkHead, 1, LSIGMA, 0, 256, 3, 1, LSIGMA};
    j[n++] = DwJob{kStashChunkBott, 4, kGradChunkZ9, 2, L9, 0, 256, 0, 128, L9};
    j[n++] = DwJob{kStashChunkEncDir, 1, kGradChunkZ9, 2, L9, 256, 27, 0, 128, -1};
    j[n++] = DwJob{kStashChunkY9, 2, kGradChunkHead, 1, LRGB, 0, 128, 0, 3, LRGB};
}

constexpr int kDwStageRows = 64;                       // rows (= UMMA K) per pipeline stage: 4 K-steps of 16
constexpr int kDwHalfChunk = kChunkBytes / 2;          // 64 rows of a chunk image = 8 KB
constexpr int kDwStageBytes = 8 * kDwHalfChunk;        // up to 4 X half-chunks + 4 dZ half-chunks = 64 KB
constexpr int kDwRingBytes = 3 * kDwStageBytes;        // 192 KB ring: 3 stages of a full job, up to 8 of a small one --
constexpr int kDwMaxStages = 8;                        // the bytes in flight per CTA, not the stage count, are constant
constexpr int kDwSmemBar = kDwRingBytes;
constexpr int kDwSmemBias = kDwSmemBar + 256;          // 256 fp32 column sums
constexpr int kDwSmemTotal = kDwSmemBias + 4096;          // [4][256] fp32 column sums
constexpr int kDwThreads = 192;                        // warps 0-3: bias + flush, warp 4: producer, warp 5: MMA

// ONE launch runs all 14 jobs of a model side by side: job j owns CTAs [cta0[j], cta0[j+1]) (its K-split,
// sized in proportion to the bytes the job streams), so the 148 CTAs produce 148 accumulator dumps per model
// instead of 148 per job, and one reduce launch folds them.
struct DwParams {
    DwJob job[kDwJobs];
    int cta0[kDwJobs + 1];
    long long poff[kDwJobs];     // float offset of job j's accumulator dumps: [split][mblocks*128][N]
    long long boff[kDwJobs];     // float offset of job j's bias column sums: [split][256]
    const uint8_t* stash; const uint8_t* gstash;
    float* partial;
    int num_tiles;
};

// MN-major SWIZZLE_128B descriptor: 64-element (128 B) atoms along MN `lbo` bytes apart, 8-row groups along K 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
__host__ __device__ constexpr uint32_t umma_idesc_mn(int fmt, int N) {   // both operands MN-major, M = 128
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

template <bool kHalf>
__global__ void __launch_bounds__(kDwThreads, 1) dw_kernel(const DwParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t sbar = sbase + kDwSmemBar;   // full[8], empty[8], done
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kDwSmemBar + 8 * (2 * kDwMaxStages + 1));
    int jidx = 0;
    while (jidx + 1 < kDwJobs && (int)blockIdx.x >= p.cta0[jidx + 1]) ++jidx;
    const DwJob jb = p.job[jidx];
    const int split = (int)blockIdx.x - p.cta0[jidx], nsplit = p.cta0[jidx + 1] - p.cta0[jidx];
    const int xn = jb.x_nchunks == 1 ? 2 : jb.x_nchunks;      // a single X chunk is loaded twice to fill M = 128
    const int mblocks = xn / 2;
    const int N = jb.dz_nchunks * 64;
    const uint32_t stage_tx = (uint32_t)(xn + jb.dz_nchunks) * kDwHalfChunk;     // = the stage stride in the ring
    const int ring = min(kDwMaxStages, kDwRingBytes / (int)stage_tx);
    auto full_bar = [&](uint32_t s) { return sbar + 8 * s; };
    auto empty_bar = [&](uint32_t s) { return sbar + 8 * (kDwMaxStages + s); };
    const uint32_t done_bar = sbar + 8 * (2 * kDwMaxStages);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kDwMaxStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1 + 4); }
        mbar_init(done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int my_tiles = split < p.num_tiles ? (p.num_tiles - 1 - split) / nsplit + 1 : 0;
    const int nstages = my_tiles * 2;           // two 64-row stages per tile
    constexpr int fmt = kHalf ? 0 : 1;

    if (warp == 4) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int it = 0; it < nstages; ++it) {
                const int tile = split + (it >> 1) * nsplit;
                const uint32_t half = (uint32_t)(it & 1) * kDwHalfChunk;
                const uint8_t* xs = p.stash + (size_t)tile * kStashTileBytes + (size_t)jb.x_chunk0 * kChunkBytes + half;
                const uint8_t* zs = p.gstash + (size_t)tile * kGradTileBytes + (size_t)jb.dz_chunk0 * kChunkBytes + half;
                mbar_wait(empty_bar(stage), phase ^ 1);
                mbar_expect_tx(full_bar(stage), stage_tx);
                const uint32_t dst = sbase + stage * stage_tx;
                for (int c = 0; c < xn; ++c)
                    bulk_g2s(dst + c * kDwHalfChunk, xs + (size_t)(jb.x_nchunks == 1 ? 0 : c) * kChunkBytes, kDwHalfChunk, full_bar(stage));
                for (int c = 0; c < jb.dz_nchunks; ++c)
                    bulk_g2s(dst + (xn + c) * kDwHalfChunk, zs + (size_t)c * kChunkBytes, kDwHalfChunk, full_bar(stage));
                if (++stage == (uint32_t)ring) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 5) {
        {   // converged warp, one elected lane issues the tcgen05 ops
            const uint32_t idesc = umma_idesc_mn(fmt, N);
            uint32_t stage = 0, phase = 0;
            for (int it = 0; it < nstages; ++it) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t xb = sbase + stage * stage_tx, zb = xb + xn * kDwHalfChunk;
                if (elect_one_sync()) {
#pragma unroll 1
                    for (int mb = 0; mb < mblocks; ++mb) {
#pragma unroll
                        for (int k = 0; k < kDwStageRows / 16; ++k) {
                            // K-step k covers rows [16k, 16k+16): two 8-row groups = 2048 B further into every atom
                            const uint64_t ad = umma_desc_mn(xb + mb * 2 * kDwHalfChunk + k * 2048, kDwHalfChunk);
                            const uint64_t bd = umma_desc_mn(zb + k * 2048, kDwHalfChunk);
                            umma_f16(tmem_base + (uint32_t)(mb * 256), ad, bd, idesc, (it == 0 && k == 0) ? 0u : 1u);
                        }
                    }
                    umma_commit(empty_bar(stage));
                    if (it == nstages - 1) umma_commit(done_bar);
                }
                __syncwarp();
                if (++stage == (uint32_t)ring) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ---- warps 0-3: bias gradient = column sums of dZ over this CTA's rows, read from the same smem stages.
        // Thread = one 16-byte unit (8 columns) x every 4th row: 16 LDS.128 per stage instead of 128 scalar loads
        // (the scalar version made this the slowest role of the CTA and capped the kernel at ~60 % of HBM).
        const int tid = threadIdx.x;             // 0..127
        const int unit = tid & 31, rg = tid >> 5;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        uint32_t stage = 0, phase = 0;
        const bool want_bias = jb.bias_layer >= 0;
        const bool my_cols = want_bias && unit * 8 < N;
        for (int it = 0; it < nstages; ++it) {
            mbar_wait(full_bar(stage), phase);
            if (my_cols) {
                const uint8_t* zb = smem + stage * stage_tx + xn * kDwHalfChunk + (unit >> 3) * kDwHalfChunk;
#pragma unroll 4
                for (int r = rg; r < kDwStageRows; r += 4) {
                    const uint4 q = *reinterpret_cast<const uint4*>(zb + r * 128 + (((unit & 7) ^ (r & 7)) << 4));
                    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float lo, hi;
                        if (kHalf) {
                            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
                            lo = f.x; hi = f.y;
                        } else {
                            lo = __uint_as_float(w[i] << 16); hi = __uint_as_float(w[i] & 0xFFFF0000u);
                        }
                        acc[2 * i] += lo;
                        acc[2 * i + 1] += hi;
                    }
                }
            }
            if (lane == 0) mbar_arrive(empty_bar(stage));     // one arrive per warp (count 1 + 4)
            __syncwarp();
            if (++stage == (uint32_t)ring) { stage = 0; phase ^= 1; }
        }
        if (want_bias) {
            // fold the 4 row groups (one per warp) in a fixed order
            float* s_red = reinterpret_cast<float*>(smem + kDwSmemBias);      // [4][256]
#pragma unroll
            for (int i = 0; i < 8; ++i) s_red[rg * 256 + unit * 8 + i] = acc[i];
            named_bar_sync(1, 128);
            float* bp = p.partial + p.boff[jidx] + (size_t)split * 256;
            for (int c = tid; c < 256; c += 128)
                bp[c] = c < N ? (s_red[c] + s_red[256 + c]) + (s_red[512 + c] + s_red[768 + c]) : 0.f;
        }
        // ---- flush the accumulators: thread = TMEM lane = X feature within the M-block
        if (nstages > 0) {
            mbar_wait(done_bar, 0);
            tc_fence_after();
        }
        float* dst = p.partial + p.poff[jidx] + (size_t)split * (size_t)(mblocks * 128) * N;
        for (int mb = 0; mb < mblocks; ++mb) {
            float* drow = dst + ((size_t)mb * 128 + warp * 32 + lane) * N;
            for (int c0 = 0; c0 < N; c0 += 32) {
                uint32_t r[32];
                if (nstages > 0) {
                    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mb * 256 + c0), r);
                    tmem_ld_wait(r);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) r[i] = 0u;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    *reinterpret_cast<uint4*>(drow + c0 + 4 * i) = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

// 3. fold the accumulator dumps of all jobs into the flat gradient buffer (+=): blockIdx.y = job, a fixed
// summation order over the job's K-splits (deterministic)
struct ReduceParams {
    DwJob job[kDwJobs];
    int nsplit[kDwJobs];
    long long poff[kDwJobs], boff[kDwJobs];
};
// sum of one element over a job's K-splits, in split order (deterministic); the loads of up to eight splits are issued
// together -- one load per iteration of a plain loop leaves a single request in flight per thread, and the launch
// is then bound by 13 dependent trips to HBM (14.5 us for 39 MB) instead of by bandwidth
__device__ __forceinline__ float fold_splits(const float* __restrict__ src, int nsplit, size_t stride) {
    float s = 0.f;
    int c = 0;
    for (; c + 8 <= nsplit; c += 8) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = __ldg(src + (size_t)(c + q) * stride);
#pragma unroll
        for (int q = 0; q < 8; ++q) s += v[q];
    }
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = (c + q < nsplit) ? __ldg(src + (size_t)(c + q) * stride) : 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q)
        if (c + q < nsplit) s += v[q];
    return s;
}
__global__ void reduce_grads_kernel(const ReduceParams rp, const float* __restrict__ partial, float* __restrict__ G /* one model */) {
    const DwJob jb = rp.job[blockIdx.y];
    const int nsplit = rp.nsplit[blockIdx.y];
    const int fan_out = layer_dim(jb.layer).fan_out;
    const int total = jb.k_rows * jb.n_cols;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int xn = jb.x_nchunks == 1 ? 2 : jb.x_nchunks;
    const int N = jb.dz_nchunks * 64;
    const size_t per_cta = (size_t)(xn / 2) * 128 * N;
    if (i < total) {
        const int m = i / jb.n_cols, n = i - m * jb.n_cols;
        const float* src = partial + rp.poff[blockIdx.y] + (size_t)m * N + jb.n_col0 + n;
        G[kernel_offset(jb.layer) + (size_t)(jb.k_row0 + m) * fan_out + n] += fold_splits(src, nsplit, per_cta);
    } else if (jb.bias_layer >= 0 && i < total + jb.n_cols) {
        const int n = i - total;
        const float* src = partial + rp.boff[blockIdx.y] + jb.n_col0 + n;
        G[bias_offset(jb.bias_layer) + n] += fold_splits(src, nsplit, 256);
    }
}

// =============================================================================================
// host side
int tc_train_create(nerfb200_ctx* ctx) {
    const BwdTable& t = bwd_table();
    for (int pz = 0; pz < 2; ++pz)
        for (int m = 0; m < 2; ++m) NB_CUDA(cudaMalloc(&ctx->packed_bwd[pz][m], (size_t)t.bytes));
    // __constant__ memory is per device: uploaded by every create, on the context's device
    NB_CUDA(cudaMemcpyToSymbol(c_bwd_chunks, t.c, sizeof(BwdChunk) * kBwdMaxChunks));
    NB_CUDA(cudaMemcpyToSymbol(c_bwd_job_begin, t.job_begin, sizeof(int) * (kBwdJobs + 1)));
    NB_CUDA(cudaFuncSetAttribute(bwd_data_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBSmemTotal));
    NB_CUDA(cudaFuncSetAttribute(bwd_data_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBSmemTotal));
    NB_CUDA(cudaFuncSetAttribute(dw_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmemTotal));
    NB_CUDA(cudaFuncSetAttribute(dw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmemTotal));
    return 0;
}

void tc_train_destroy(nerfb200_ctx* ctx) {
    for (int pz = 0; pz < 2; ++pz)
        for (int m = 0; m < 2; ++m) if (ctx->packed_bwd[pz][m]) cudaFree(ctx->packed_bwd[pz][m]);
}

int tc_train_pack(nerfb200_ctx* ctx, const float* flat_params, cudaStream_t st) {
    const BwdTable& t = bwd_table();
    const int64_t units = (int64_t)t.n * (kChunkBytes / 16);
    const dim3 grid((unsigned)((units + 255) / 256), 2);          // both models in one launch
    if (ctx->pack_mask & 1) {
        pack_bwd_weights_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(flat_params, (uint8_t*)ctx->packed_bwd[0][0], (uint8_t*)ctx->packed_bwd[0][1], t.n);
        NB_LAUNCH_CHECK();
    }
    if (ctx->pack_mask & 2) {
        pack_bwd_weights_kernel<__half><<<grid, 256, 0, st>>>(flat_params, (uint8_t*)ctx->packed_bwd[1][0], (uint8_t*)ctx->packed_bwd[1][1], t.n);
        NB_LAUNCH_CHECK();
    }
    return 0;
}

static inline int64_t tiles_of(int64_t R) { return (R + kTileRows - 1) / kTileRows; }

int64_t tc_stash_bytes(int64_t R) { return tiles_of(R) * (int64_t)kStashTileBytes; }

// backward workspace: gradient stash + one accumulator dump (<= 256 x 256 fp32) and 256 bias sums per CTA
static int64_t dw_partial_floats(int grid) { return (int64_t)grid * (256 * 256 + 256); }
int64_t tc_workspace_bytes(int64_t R, int training) {
    if (!training) return 0;
    const int grid = num_sms() > kDwJobs ? num_sms() : kDwJobs;
    return tiles_of(R) * (int64_t)kGradTileBytes + dw_partial_floats(grid) * 4;
}

// K-split sizes: every job gets one CTA, the rest go one at a time to the job with the most bytes per CTA
static void plan_dw(const DwJob* jobs, int grid, DwParams& dp, ReduceParams& rp) {
    int cost[kDwJobs], nsplit[kDwJobs];
    for (int j = 0; j < kDwJobs; ++j) {
        cost[j] = (jobs[j].x_nchunks == 1 ? 2 : jobs[j].x_nchunks) + jobs[j].dz_nchunks;
        nsplit[j] = 1;
    }
    for (int c = kDwJobs; c < grid; ++c) {
        int best = 0;
        for (int j = 1; j < kDwJobs; ++j)
            if ((long long)cost[j] * nsplit[best] > (long long)cost[best] * nsplit[j]) best = j;
        ++nsplit[best];
    }
    long long off = 0;
    dp.cta0[0] = 0;
    for (int j = 0; j < kDwJobs; ++j) {
        const int xn = jobs[j].x_nchunks == 1 ? 2 : jobs[j].x_nchunks;
        dp.job[j] = rp.job[j] = jobs[j];
        rp.nsplit[j] = nsplit[j];
        dp.cta0[j + 1] = dp.cta0[j] + nsplit[j];
        dp.poff[j] = rp.poff[j] = off;
        off += (long long)nsplit[j] * (xn / 2) * 128 * jobs[j].dz_nchunks * 64;
        dp.boff[j] = rp.boff[j] = off;
        off += (long long)nsplit[j] * 256;
    }
}

// The backward pass of one model is two phases with different HBM behaviour: backward-data WRITES the gradient stash
// (write bound), the weight-gradient GEMM READS both stashes (read bound). `max_sms` > 0 caps the SMs a phase may
// occupy (both kernels are persistent, one CTA per SM), so that the caller can run the coarse model's weight-gradient
// phase next to the fine model's backward-data phase on disjoint SMs and keep reads and writes in flight together.
int tc_backward_data(nerfb200_ctx* ctx, int which, int half, int64_t B, int S, const float* flat_params, const float* d_rgb,
                     const float* d_sigma, void* workspace, void* stash, int max_sms, cudaStream_t st) {
    int rc = check_device(ctx, "mlp_backward");
    if (rc) return rc;
    if (!ctx->packed_valid || !(ctx->packed_mask & (half ? 2 : 1))) {
        set_error("mlp_backward: pack_weights has not been called for this precision");
        return NERFB200_ESTATE;
    }
    NB_CHECK_ARG(workspace && stash, "mlp_backward: workspace and stash required");
    const int64_t R = B * S;
    if (R == 0) return 0;
    const int num_tiles = (int)tiles_of(R);
    int sms = ctx->num_sms;
    if (max_sms > 0 && max_sms < sms) sms = max_sms < 2 ? 2 : max_sms;
    uint8_t* gstash = (uint8_t*)workspace;
    const float* P = flat_params + (int64_t)which * kParamsPerModel;

    BwdParams bp;
    bp.wimg = (const uint8_t*)ctx->packed_bwd[half ? 1 : 0][which];
    bp.P = P; bp.stash = (const uint8_t*)stash; bp.gstash = gstash; bp.d_rgb = d_rgb; bp.d_sigma = d_sigma; bp.R = R; bp.num_tiles = num_tiles;
    const int quads = (num_tiles + 3) / 4;
    const int clusters = quads < sms / 2 ? quads : sms / 2;
    if (half) bwd_data_pair_kernel<true><<<2 * clusters, kBThreads, kBSmemTotal, st>>>(bp);
    else bwd_data_pair_kernel<false><<<2 * clusters, kBThreads, kBSmemTotal, st>>>(bp);
    NB_LAUNCH_CHECK();
    return 0;
}

int tc_backward_weights(nerfb200_ctx* ctx, int which, int half, int64_t B, int S, float* flat_grads, void* workspace,
                        void* stash, int max_sms, cudaStream_t st) {
    NB_CHECK_ARG(workspace && stash, "mlp_backward: workspace and stash required");
    int rc = check_device(ctx, "mlp_backward_weights");
    if (rc) return rc;
    const int64_t R = B * S;
    if (R == 0) return 0;
    const int num_tiles = (int)tiles_of(R);
    int sms = ctx->num_sms;
    if (max_sms > 0 && max_sms < sms) sms = max_sms;
    uint8_t* gstash = (uint8_t*)workspace;
    float* partial0 = (float*)(gstash + (size_t)num_tiles * kGradTileBytes);
    float* G = flat_grads + (int64_t)which * kParamsPerModel;

    DwJob jobs[kDwJobs];
    build_dw_jobs(jobs);
    int dgrid = sms > kDwJobs ? sms : kDwJobs;
    const bool capped = max_sms > 0 && max_sms < ctx->num_sms;
    if (capped) dgrid &= ~1;
    DwParams dp;
    ReduceParams rp;
    plan_dw(jobs, dgrid, dp, rp);
    dp.stash = (const uint8_t*)stash; dp.gstash = gstash; dp.partial = partial0; dp.num_tiles = num_tiles;
    if (capped) {
        // launched as 2-CTA clusters (the kernel itself does not use the cluster): a pair lands on the two SMs of one
        // TPC, so the SMs this phase leaves free are whole TPCs, which is what the cta_group::2 pairs of the
        // backward-data kernel running next to it need
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)dgrid); cfg.blockDim = dim3(kDwThreads); cfg.dynamicSmemBytes = kDwSmemTotal; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (half) NB_CUDA(cudaLaunchKernelEx(&cfg, dw_kernel<true>, dp));
        else NB_CUDA(cudaLaunchKernelEx(&cfg, dw_kernel<false>, dp));
    } else if (half) dw_kernel<true><<<dgrid, kDwThreads, kDwSmemTotal, st>>>(dp);
    else dw_kernel<false><<<dgrid, kDwThreads, kDwSmemTotal, st>>>(dp);
    NB_LAUNCH_CHECK();
    reduce_grads_kernel<<<dim3((256 * 256 + 256 + 255) / 256, kDwJobs), 256, 0, st>>>(rp, partial0, G);
    NB_LAUNCH_CHECK();
    return 0;
}

int tc_backward(nerfb200_ctx* ctx, int which, int half, int64_t B, int S, const float* ro, const float* rd, const float* t,
                const float* flat_params, const float* d_rgb, const float* d_sigma, float* flat_grads, void* workspace, void* stash,
                cudaStream_t st) {
    (void)ro; (void)rd; (void)t;
    int rc = tc_backward_data(ctx, which, half, B, S, flat_params, d_rgb, d_sigma, workspace, stash, 0, st);
    if (rc) return rc;
    return tc_backward_weights(ctx, which, half, B, S, flat_grads, workspace, stash, 0, st);
}

}  // namespace nb
