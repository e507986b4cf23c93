// Fused positional-encoding + 8x256 MLP forward (core/model.py:289-394) with tcgen05.mma.kind::tf32:
// the reference's own arithmetic width on Ampere-and-later GPUs, where TensorFlow runs fp32 MatMul as TF32
// (SURVEY.md App. B11). Render path only (training uses the 16-bit kernels of mlp_tc.cu / mlp_tc_train.cu).
//
// 32-bit operands double every shared-memory footprint of the 16-bit pair kernel: a 128-row tile's 256-wide
// activation is 128 KB, so a CTA holds ONE tile and the layer-granular ping-pong between two tiles is replaced
// by a chunk-granular pipeline inside the tile:
//   * two fp32 accumulators in TMEM (2 x 256 columns) alternate by layer;
//   * the epilogue of layer l (8 warps: two per TMEM lane quadrant, each pair splitting the columns) writes the
//     next A operand in place, 32 columns (= one [128 x 32] tf32 K chunk, 16 KB) at a time, and signals one
//     "chunk ready" barrier per chunk;
//   * the MMA warp issues layer l+1's K chunk c as soon as chunk c is ready, into the other accumulator, so
//     the tensor pipe works on layer l+1 while layer l is still being drained. In place is safe: the epilogue of
//     layer l starts after ALL of layer l's MMAs have completed (acc_full), i.e. after the last read of the buffer.
// tf32 MMAs run at half the 16-bit rate (M=256, N=256, K=8 = 128 tensor cycles), so a layer is 32 instructions =
// 4096 cycles of tensor work; what bounds the kernel is shared-memory bandwidth (every operand is twice as wide as in
// the 16-bit kernel: >= 512 KB per layer and SM through a 128 B/clk pipe), so everything that can stay out of shared
// memory does: the layer bias is added by the tensor core (a "ones" x bias-tile MMA starts each accumulator, as in
// the pair kernel) and the sigma head's weights are read through L1.
//
// 2-CTA clusters, cta_group::2 (M = 256 = one tile per CTA), each CTA streams HALF of every weight chunk through
// a 4 x 16 KB bulk-TMA ring; roles: warps 0-7 epilogue, warp 8 weight producer, warp 9 MMA issuer (leader CTA) /
// relay of "my half landed" (peer CTA).
//
// Jobs of one tile (K chunks of 32; issue order = encoding chunks first, they are ready long before):
//   J0  dense_0   enc_xyz 2 chunks (63 + 1 zero column)          N=256
//   J1-4,6-8      act 8 chunks                                    N=256   (J7: + sigma head in the epilogue, fp32)
//   J5  dense_5   enc_xyz 2 chunks + act 8 chunks (skip concat)   N=256
//   J9  dense_9   enc_dir 1 chunk (27 + 5 zero columns) + act 8   N=128
//   J10 rgb       act 4 chunks                                    N=16 (3 used), sigmoid in the epilogue
#include "common.cuh"
#include "mlp.cuh"
#include "tc_common.cuh"

namespace nb {

namespace {

constexpr int kJobs = 11;
constexpr int kChunk = 16384;            // [128 rows x 32 K] fp32, 128 B per row, 16-byte units XOR-swizzled by (row & 7)
constexpr int kStagesT = 4;
constexpr int kSmemActT = 0;                               // 8 chunks = 128 KB
constexpr int kSmemEncT = kSmemActT + 8 * kChunk;          // 2 chunks = 32 KB (enc_xyz; enc_dir re-uses chunk 0)
constexpr int kSmemRingT = kSmemEncT + 2 * kChunk;         // 4 x 16 KB
constexpr int kSmemBarT = kSmemRingT + kStagesT * kChunk;
constexpr int kHeadWsig = 9 * 256 + 128 + 16;              // float offsets inside the fp32 head block (HeadOffsets of mlp_tc.cu):
constexpr int kHeadBsig = kHeadWsig + 256;                 //   sigma kernel (256), sigma bias
constexpr int kSmemOnesT = kSmemBarT + 256;                // 256 B "ones" A operand of the bias MMA step
constexpr int kSmemBiasTileT = kSmemOnesT + 256;           // one job's bias tile (this CTA's N/2 rows x 16 B)
constexpr int kBiasTileBytesT = 2048;
constexpr int kSmemSigT = kSmemBiasTileT + kBiasTileBytesT;
constexpr int kSmemTotalT = kSmemSigT + 128 * 4;
static_assert(kSmemTotalT <= 232448, "exceeds the 227 KB dynamic shared memory limit");
constexpr int kThreadsT = 320;
constexpr int kProducerWarpT = 8, kMmaWarpT = 9;

__host__ __device__ constexpr int job_layer(int j) { return j <= 9 ? j : (int)LRGB; }      // L0..L9 = 0..9
__host__ __device__ constexpr int job_nchunks(int j) { return j == 0 ? 2 : j == 5 ? 10 : j == 9 ? 9 : j == 10 ? 4 : 8; }
__host__ __device__ constexpr int job_nenc(int j) { return (j == 0 || j == 5) ? 2 : j == 9 ? 1 : 0; }
__host__ __device__ constexpr int job_rows(int j) { return j < 9 ? 256 : j == 9 ? 128 : 16; }     // N
// image: [job][issue-order chunk][N rows x 128 B]; a CTA's half of a chunk is rows [rank * N/2, (rank+1) * N/2)
__host__ __device__ constexpr uint32_t job_ofs(int j) {
    uint32_t o = 0;
    for (int i = 0; i < j; ++i) o += (uint32_t)job_nchunks(i) * (uint32_t)job_rows(i) * 128u;
    return o;
}
constexpr uint32_t kChunksBytes = job_ofs(kJobs);
// behind the chunks: the bias tiles [job][cta rank][kBiasTileBytesT] -- the layer bias as a tensor-core B operand: N-row i
// of the CTA's share is the 16 bytes [hi, mid, lo, 0] in tf32 (hi + mid + lo = the fp32 bias exactly); the first MMA of a
// layer multiplies the "ones" A operand [1, 1, 1, 0 | 0 ...] with it (accumulate off), as in the 16-bit pair kernel
constexpr uint32_t kImageBytes = kChunksBytes + kJobs * 2 * kBiasTileBytesT;

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t d;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(d) : "f"(x));
    return d;
}

// first fan-in row of issue-order chunk i of job j in the Keras kernel [in, out], and how many rows are real
__device__ __forceinline__ void chunk_k_range(int j, int i, int& k0, int& kvalid) {
    const int ne = job_nenc(j);
    if (i < ne) {                          // encoding chunk
        if (j == 0) { k0 = 32 * i; kvalid = i == 0 ? 32 : 31; }
        else if (j == 5) { k0 = 256 + 32 * i; kvalid = i == 0 ? 32 : 31; }
        else { k0 = 256; kvalid = 27; }    // j == 9: enc_dir
    } else { k0 = 32 * (i - ne); kvalid = 32; }
}

// One thread per 16-byte unit (4 consecutive K of one N row) of the tf32 image.
__global__ void pack_tf32_kernel(const float* __restrict__ P /* one model */, uint8_t* __restrict__ img) {
    const uint32_t byte = (uint32_t)(blockIdx.x * blockDim.x + threadIdx.x) * 16u;
    if (byte >= kImageBytes) return;
    if (byte >= kChunksBytes) {          // bias tile row: (job, rank, row)
        const uint32_t i = (byte - kChunksBytes) >> 4;
        const int job = (int)(i >> 8), rnk = (int)((i >> 7) & 1), row = (int)(i & 127);
        const int rows = job_rows(job) / 2, n = rnk * rows + row, l = job_layer(job);
        uint32_t v[4] = {0u, 0u, 0u, 0u};
        if (row < rows && n < layer_dim(l).fan_out) {
            const float b = P[bias_offset(l) + n];
            v[0] = to_tf32(b);
            const float r1 = b - __uint_as_float(v[0]);
            v[1] = to_tf32(r1);
            v[2] = to_tf32(r1 - __uint_as_float(v[1]));
        }
        *reinterpret_cast<uint4*>(img + byte) = make_uint4(v[0], v[1], v[2], v[3]);
        return;
    }
    int j = 0;
#pragma unroll 1
    while (j + 1 < kJobs && byte >= job_ofs(j + 1)) ++j;
    const uint32_t local = byte - job_ofs(j);
    const uint32_t chunk_bytes = (uint32_t)job_rows(j) * 128u;
    const int i = (int)(local / chunk_bytes);
    const uint32_t in_chunk = local % chunk_bytes;
    const int n = (int)(in_chunk >> 7);
    const int unit = (int)((in_chunk >> 4) & 7) ^ (n & 7);
    int k0, kvalid;
    chunk_k_range(j, i, k0, kvalid);
    const int l = job_layer(j);
    const LayerDim dim = layer_dim(l);
    const float* W = P + kernel_offset(l);
    uint32_t v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int kl = unit * 4 + e;
        float w = 0.f;
        if (kl < kvalid && n < dim.fan_out) w = W[(int64_t)(k0 + kl) * dim.fan_out + n];
        v[e] = to_tf32(w);
    }
    *reinterpret_cast<uint4*>(img + byte) = make_uint4(v[0], v[1], v[2], v[3]);
}

struct Tf32Params {
    const uint8_t* wimg;
    const float* heads;      // HeadOffsets block of mlp_tc.cu (same layout as kHead*)
    const float* ro; const float* rd; const float* t;
    float* rgb; float* sigma;
    int64_t R;
    int S;
    int num_tiles;
};

struct RowT { int64_t grow; bool valid; float dir[3]; };

// sin/cos of the L octaves of one coordinate: the reference's arguments fl32(x * fl32(2^l pi)) are exactly 2^l * a0
// with a0 = fl32(x * fl32(pi)), so an accurate sincosf every 5th octave + double-angle steps evaluate the SAME
// arguments to ~3e-6, two orders below the tf32 rounding of the operand (see mlp_tc.cu sincos_octaves).
template <int L>
__device__ __forceinline__ void octaves(float x, float* e) {
    const float a0 = __fmul_rn(x, 3.14159274101257324f);
    float sn = 0.f, cs = 1.f;
#pragma unroll
    for (int l = 0; l < L; ++l) {
        if (l % 5 == 0) sincosf(__fmul_rn(a0, (float)(1 << l)), &sn, &cs);
        else {
            const float s2 = __fmul_rn(__fmul_rn(2.f, sn), cs);
            cs = fmaf(__fmul_rn(-2.f, sn), sn, 1.f);
            sn = s2;
        }
        e[2 * l] = sn;
        e[2 * l + 1] = cs;
    }
}

__device__ __forceinline__ void store_chunk_row(uint8_t* chunk, int row, const float* e) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
        *reinterpret_cast<uint4*>(chunk + swz(row, u)) =
            make_uint4(to_tf32(e[4 * u]), to_tf32(e[4 * u + 1]), to_tf32(e[4 * u + 2]), to_tf32(e[4 * u + 3]));
}

// 32 accumulator columns of this thread's row (the bias is already in the accumulator) -> (ReLU) -> tf32 -> one swizzled
// chunk row
template <bool kRelu, bool kSigma>
__device__ __forceinline__ void drain32(const uint32_t (&rr)[32], uint8_t* chunk, int row, const float* wsig, float (&sg)[4]) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        float v0 = __uint_as_float(rr[4 * u + 0]), v1 = __uint_as_float(rr[4 * u + 1]);
        float v2 = __uint_as_float(rr[4 * u + 2]), v3 = __uint_as_float(rr[4 * u + 3]);
        if (kRelu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
        if (kSigma) {      // sigma head on the fp32 activations (core/model.py:375)
            const float4 w = __ldg(reinterpret_cast<const float4*>(wsig) + u);
            sg[0] = fmaf(v0, w.x, sg[0]); sg[1] = fmaf(v1, w.y, sg[1]); sg[2] = fmaf(v2, w.z, sg[2]); sg[3] = fmaf(v3, w.w, sg[3]);
        }
        *reinterpret_cast<uint4*>(chunk + swz(row, u)) = make_uint4(to_tf32(v0), to_tf32(v1), to_tf32(v2), to_tf32(v3));
    }
}

__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsT, 1) mlp_tf32_forward_kernel(const Tf32Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t sbase = smem_u32(smem);
    const uint32_t sbar = sbase + kSmemBarT;
    auto ring_full = [&](int s) { return sbar + 8 * s; };
    auto ring_empty = [&](int s) { return sbar + 8 * (kStagesT + s); };
    auto chunk_ready = [&](int c) { return sbar + 8 * (2 * kStagesT + c); };          // leader: 4 warps x 2 CTAs
    const uint32_t enc_ready = sbar + 8 * (2 * kStagesT + 8);                         // leader: 8 warps x 2 CTAs
    auto acc_full = [&](int b) { return sbar + 8 * (2 * kStagesT + 9 + b); };         // both CTAs (multicast commit)
    const uint32_t bias_full = sbar + 8 * (2 * kStagesT + 11), bias_empty = sbar + 8 * (2 * kStagesT + 12);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kSmemBarT + 8 * (2 * kStagesT + 13));
    float* s_sig = reinterpret_cast<float*>(smem + kSmemSigT);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesT; ++s) { mbar_init(ring_full(s), rank == 0 ? 2 : 1); mbar_init(ring_empty(s), 1); }
        for (int c = 0; c < 8; ++c) mbar_init(chunk_ready(c), 8);
        mbar_init(enc_ready, 16);
        mbar_init(acc_full(0), 1); mbar_init(acc_full(1), 1);
        mbar_init(bias_full, rank == 0 ? 2 : 1);     // like ring_full: own expect_tx arrive (+ the peer's relay in the leader)
        mbar_init(bias_empty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 16) {
        // "ones" A operand of the bias step: core matrix 0 = 8 rows x [1, 1, 1, 0] (tf32), core matrix 1 = zeros; its
        // descriptor has SBO = 0, so all sixteen 8-row groups of the 128-row tile read these same 256 bytes
        const uint32_t one = __float_as_uint(1.0f);
        reinterpret_cast<uint4*>(smem + kSmemOnesT)[threadIdx.x] = threadIdx.x < 8 ? make_uint4(one, one, one, 0u) : make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async();
    }
    cluster_sync_all();
    if (warp == kMmaWarpT) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int num_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
    const int tpairs = (p.num_tiles + 1) >> 1;          // a cluster works on 2 tiles at a time (one per CTA)

    if (warp == kProducerWarpT) {
        // ===================== weight producer: this CTA's half of every chunk =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, bphase = 0;
            for (int tp = cluster_id; tp < tpairs; tp += num_clusters)
#pragma unroll 1
                for (int j = 0; j < kJobs; ++j) {
                    {   // this CTA's share of the layer's bias tile (single buffer, released by the bias MMA's commit)
                        const uint32_t bbytes = (uint32_t)job_rows(j) * 8u;
                        mbar_wait(bias_empty, bphase ^ 1);
                        mbar_expect_tx(bias_full, bbytes);
                        bulk_g2s(sbase + kSmemBiasTileT, p.wimg + kChunksBytes + (uint32_t)(j * 2 + (int)rank) * kBiasTileBytesT, bbytes, bias_full);
                        bphase ^= 1;
                    }
                    const uint32_t half_bytes = (uint32_t)job_rows(j) * 64u;
                    const uint8_t* src = p.wimg + job_ofs(j) + rank * half_bytes;
                    const int NC = job_nchunks(j);
#pragma unroll 1
                    for (int i = 0; i < NC; ++i) {
                        mbar_wait(ring_empty(stage), phase ^ 1);
                        mbar_expect_tx(ring_full(stage), half_bytes);
                        bulk_g2s(sbase + kSmemRingT + stage * kChunk, src + (size_t)i * 2u * half_bytes, half_bytes, ring_full(stage));
                        if (++stage == kStagesT) { stage = 0; phase ^= 1; }
                    }
                }
        }
    } else if (warp == kMmaWarpT) {
        if (rank == 1) {
            // ===================== peer: relay "my half of the stage has landed" to the leader =====================
            if (lane == 0) {
                uint32_t stage = 0, phase = 0, bphase = 0;
                const uint32_t bias_full_leader = mapa(bias_full, 0);
                for (int tp = cluster_id; tp < tpairs; tp += num_clusters)
#pragma unroll 1
                    for (int j = 0; j < kJobs; ++j) {
                        mbar_wait(bias_full, bphase);
                        mbar_arrive_cluster(bias_full_leader);
                        bphase ^= 1;
                        for (int i = 0; i < job_nchunks(j); ++i) {
                            mbar_wait(ring_full(stage), phase);
                            mbar_arrive_cluster(mapa(ring_full(stage), 0));
                            if (++stage == kStagesT) { stage = 0; phase ^= 1; }
                        }
                    }
            }
        } else {
            // ===================== leader: MMA issuer for the pair =====================
            // the whole warp walks the loop converged; ONE election per job, the elected lane walks the chunks alone
            const uint32_t ring_lo = ((sbase + kSmemRingT) >> 4) & 0x3FFFu;
            const uint32_t act_lo = ((sbase + kSmemActT) >> 4) & 0x3FFFu, enc_lo = ((sbase + kSmemEncT) >> 4) & 0x3FFFu;
            constexpr uint32_t id256 = umma_idesc_pair(2, 256), id128 = umma_idesc_pair(2, 128), id16 = umma_idesc_pair(2, 16);
            uint32_t stage = 0, phase = 0, chunk_phase_bits = 0, enc_phase = 0, bphase = 0;
            // no-swizzle K-major descriptors of the bias step: A = the ones atom (LBO 128 B between its two core matrices,
            // SBO 0: every 8-row group aliases it), B = the bias tile (SBO 128 B between 8-row groups, LBO 0: k 4..7 alias
            // k 0..3 and meet A's zeros)
            const uint64_t ones_desc = (uint64_t)(((sbase + kSmemOnesT) >> 4) & 0x3FFFu) | ((uint64_t)(128u >> 4) << 16) | (1ull << 46);
            const uint64_t biast_desc = (uint64_t)(((sbase + kSmemBiasTileT) >> 4) & 0x3FFFu) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
            for (int tp = cluster_id; tp < tpairs; tp += num_clusters) {
#pragma unroll 1
                for (int j = 0; j < kJobs; ++j) {
                    const int NC = job_nchunks(j), NE = job_nenc(j);
                    const uint32_t idesc = (j < 9) ? id256 : (j == 9) ? id128 : id16;
                    const int buf = (j == 10) ? 1 : (j & 1);
                    const uint32_t d = tmem_base + (uint32_t)(buf * 256);
                    const bool wait_enc = (j == 0 || j == 9);         // a freshly written encoding buffer
                    if (elect_one_sync()) {
                        // the layer starts from its bias: D = ones . bias_tile^T (accumulate off). The accumulator is free:
                        // the previous layer's MMAs, all issued, waited for every chunk of the epilogue that last read it.
                        mbar_wait_cluster(bias_full, bphase);
                        tc_fence_after();
                        umma_tf32_pair(d, ones_desc, biast_desc, idesc, 0u);
                        umma_commit_pair(bias_empty);
                        if (wait_enc) { mbar_wait_cluster(enc_ready, enc_phase); tc_fence_after(); }
                        uint32_t st = stage, ph = phase;
#pragma unroll 1
                        for (int i = 0; i < NC; ++i) {
                            uint32_t a_lo;
                            if (i < NE) a_lo = enc_lo + (uint32_t)(i * (kChunk >> 4));
                            else {
                                const int ac = i - NE;
                                mbar_wait_cluster(chunk_ready(ac), (chunk_phase_bits >> ac) & 1u);
                                a_lo = act_lo + (uint32_t)(ac * (kChunk >> 4));
                            }
                            mbar_wait_cluster(ring_full(st), ph);
                            tc_fence_after();
                            const uint32_t b_lo = ring_lo + st * (kChunk >> 4);
#pragma unroll
                            for (int k = 0; k < 4; ++k)      // 4 x K=8 (32 bytes of a 128-byte swizzled row each)
                                umma_tf32_pair(d, umma_desc_from_lo(a_lo + 2 * k), umma_desc_from_lo(b_lo + 2 * k), idesc, 1u);
                            umma_commit_pair(ring_empty(st));
                            if (++st == kStagesT) { st = 0; ph ^= 1; }
                        }
                        umma_commit_pair(acc_full(buf));
                    }
                    __syncwarp();
                    // every lane advances the pipeline state
                    bphase ^= 1;
                    if (wait_enc) enc_phase ^= 1;
                    chunk_phase_bits ^= (1u << (NC - NE)) - 1u;
                    phase ^= ((stage + (uint32_t)NC) / kStagesT) & 1u;
                    stage = (stage + (uint32_t)NC) % kStagesT;
                }
            }
        }
    } else {
        // ===================== 8 epilogue warps: grp = warp >> 2 splits the columns, warp & 3 = TMEM lane quadrant
        const int grp = warp >> 2, q = warp & 3, row = q * 32 + lane;
        uint8_t* act = smem + kSmemActT;
        uint8_t* enc = smem + kSmemEncT;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t enc_ready_leader = mapa(enc_ready, 0);
        uint32_t acc_phase_bits = 0;
        const float* wsig = p.heads + kHeadWsig;        // fp32 sigma kernel, read through L1 (the same 1 KB for every thread)
        const float bsig = __ldg(p.heads + kHeadBsig);

        auto load_row = [&](int tile, RowT& rc, float (&xyz)[3]) {
            rc.grow = (int64_t)tile * kTileRows + row;
            rc.valid = tile < p.num_tiles && rc.grow < p.R;
            const int64_t lrow = rc.valid ? rc.grow : p.R - 1;
            const int64_t ray = (p.R <= 0x7fffffffLL) ? (int64_t)((uint32_t)lrow / (uint32_t)p.S) : lrow / p.S;
            const float tv = __ldg(p.t + lrow);
#pragma unroll
            for (int dd = 0; dd < 3; ++dd) {
                rc.dir[dd] = __ldg(p.rd + 3 * ray + dd);
                xyz[dd] = __fadd_rn(__ldg(p.ro + 3 * ray + dd), __fmul_rn(tv, rc.dir[dd]));      // utils/ray_utils.py:251
            }
        };
        // enc_xyz (core/model.py:305-332, L = 10): 63 features + one zero column = two chunks; group g writes chunk g
        auto write_enc_xyz = [&](const float (&xyz)[3]) {
            float e[64];
            e[0] = xyz[0]; e[1] = xyz[1]; e[2] = xyz[2];
#pragma unroll
            for (int dd = 0; dd < 3; ++dd) octaves<10>(xyz[dd], e + 3 + dd * 20);
            e[63] = 0.f;
            if (grp == 0) store_chunk_row(enc, row, e);
            else store_chunk_row(enc + kChunk, row, e + 32);
        };
        auto signal = [&](uint32_t cluster_bar) {      // this warp's share of an operand is written: publish it
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(cluster_bar);
        };

        RowT cur, nxt;
        cur.grow = 0; cur.valid = false; cur.dir[0] = cur.dir[1] = cur.dir[2] = 0.f;
        nxt = cur;
        int tp = cluster_id;
        if (tp < tpairs) {
            float xyz[3];
            load_row(tp * 2 + (int)rank, cur, xyz);
            write_enc_xyz(xyz);
            signal(enc_ready_leader);
        }
        for (; tp < tpairs; tp += num_clusters) {
            float sg[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
            for (int j = 0; j < kJobs; ++j) {
                const int buf = (j == 10) ? 1 : (j & 1);
                mbar_wait(acc_full(buf), (acc_phase_bits >> buf) & 1u);
                acc_phase_bits ^= 1u << buf;
                tc_fence_after();
                const uint32_t tcols = tmem_lane + (uint32_t)(buf * 256);
                if (j < 10) {
                    // The two groups take the 32-column chunks ALTERNATELY (group g: chunks g, g + 2, g + 4, ...), so that
                    // chunks become ready in the order the next layer's MMAs consume them, two at a time: with each
                    // group owning one contiguous half, chunk 1 was ready only after chunk 0 AND 1 of the same warps
                    // while chunks 4.. sat finished and unused. N = 256: 4 chunks per group; dense_9 (N = 128): 2.
                    const int nch = j == 9 ? 2 : 4;
                    uint32_t r[2][32];
                    tmem_ld32(tcols + 32u * grp, r[0]);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (c < nch) {
                            const int ch = grp + 2 * c;
                            tmem_ld_wait(r[c & 1]);
                            if (c + 1 < nch) tmem_ld32(tcols + 32u * (ch + 2), r[(c + 1) & 1]);
                            if (j == 7) drain32<true, true>(r[c & 1], act + ch * kChunk, row, wsig + 32 * ch, sg);
                            else if (j == 8) drain32<false, false>(r[c & 1], act + ch * kChunk, row, wsig, sg);
                            else drain32<true, false>(r[c & 1], act + ch * kChunk, row, wsig, sg);
                            signal(mapa(chunk_ready(ch), 0));
                        }
                    }
                    if (j == 5) {
                        // dense_5 has consumed enc_xyz: chunk 0 of the encoding buffer now takes enc_dir (27 features + zeros)
                        if (grp == 0) {
                            float e[32];
                            e[0] = cur.dir[0]; e[1] = cur.dir[1]; e[2] = cur.dir[2];
#pragma unroll
                            for (int dd = 0; dd < 3; ++dd) octaves<4>(cur.dir[dd], e + 3 + dd * 8);
#pragma unroll
                            for (int i = 27; i < 32; ++i) e[i] = 0.f;
                            store_chunk_row(enc, row, e);
                        }
                        signal(enc_ready_leader);
                    }
                    if (j == 7) {
                        const float part = (sg[0] + sg[1]) + (sg[2] + sg[3]);
                        if (grp == 1) s_sig[row] = part;
                        named_bar_sync(1, 256);
                        if (grp == 0 && cur.valid) p.sigma[cur.grow] = fmaxf(part + s_sig[row] + bsig, 0.f);
                    }
                    if (j == 9) {
                        // dense_9 has consumed enc_dir: encode the next tile now
                        const int ntp = tp + num_clusters;
                        if (ntp < tpairs) {
                            float xyz[3];
                            load_row(ntp * 2 + (int)rank, nxt, xyz);
                            write_enc_xyz(xyz);
                            signal(enc_ready_leader);
                        }
                    }
                } else {
                    if (grp == 0) {
                        uint32_t r[32];
                        tmem_ld32(tcols, r);
                        tmem_ld_wait(r);
                        if (cur.valid) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                const float x = __uint_as_float(r[c]);        // bias included by the tensor core
                                p.rgb[3 * cur.grow + c] = 1.f / (1.f + expf(-x));
                            }
                        }
                    }
                    tc_fence_before();
                }
            }
            cur = nxt;
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == kMmaWarpT) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

}  // namespace

int tf32_create(nerfb200_ctx* ctx) {
    for (int m = 0; m < 2; ++m) NB_CUDA(cudaMalloc(&ctx->packed_tf32[m], kImageBytes));
    NB_CUDA(cudaFuncSetAttribute(mlp_tf32_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotalT));
    return 0;
}

void tf32_destroy(nerfb200_ctx* ctx) {
    for (int m = 0; m < 2; ++m) if (ctx->packed_tf32[m]) cudaFree(ctx->packed_tf32[m]);
}

int tf32_pack(nerfb200_ctx* ctx, const float* flat_params, cudaStream_t st) {
    for (int m = 0; m < 2; ++m) {
        pack_tf32_kernel<<<(kImageBytes / 16 + 255) / 256, 256, 0, st>>>(flat_params + (int64_t)m * kParamsPerModel, (uint8_t*)ctx->packed_tf32[m]);
        NB_LAUNCH_CHECK();
    }
    return 0;
}

int tf32_forward(nerfb200_ctx* ctx, int which, int64_t B, int S, const float* ro, const float* rd, const float* t,
                 float* rgb, float* sigma, cudaStream_t st) {
    int rc = check_device(ctx, "mlp_forward");
    if (rc) return rc;
    if (!ctx->packed_valid || !(ctx->packed_mask & 4)) {
        set_error("mlp_forward: pack_weights has not been called for tf32");
        return NERFB200_ESTATE;
    }
    const int64_t R = B * S;
    if (R == 0) return 0;
    NB_CHECK_ARG((R + kTileRows - 1) / kTileRows < (int64_t)1 << 30, "mlp_forward: too many rows");
    Tf32Params p{};
    p.wimg = (const uint8_t*)ctx->packed_tf32[which];
    p.heads = ctx->head_params[which];
    p.ro = ro; p.rd = rd; p.t = t; p.rgb = rgb; p.sigma = sigma; p.R = R; p.S = S;
    p.num_tiles = (int)((R + kTileRows - 1) / kTileRows);
    const int tpairs = (p.num_tiles + 1) / 2;
    const int clusters = tpairs < ctx->num_sms / 2 ? tpairs : ctx->num_sms / 2;
    mlp_tf32_forward_kernel<<<2 * clusters, kThreadsT, kSmemTotalT, st>>>(p);
    NB_LAUNCH_CHECK();
    // the last sample of every ray with split bf16 operands (more significand bits than tf32; see mlp_tc.cu)
    return tc_precise_last(ctx, which, 0, B, S, ro, rd, t, sigma, nullptr, st);
}

}  // namespace nb
