// Fused positional-encoding + 8x256 MLP forward (core/model.py:289-394) on the 5th-generation
// tensor cores: tcgen05.mma with fp32 accumulators in TMEM, 16-bit (bf16 or fp16) operands in
// 128B-swizzled shared memory, weights streamed by bulk-TMA (cp.async.bulk, UBLKCP) through an
// mbarrier ring. One persistent CTA per SM; per-sample activations never leave the SM.
//
// CTA layout (320 threads):
//   warps 0-3   : epilogue of row-tile slot 0 (thread = row = TMEM lane; warp%4 = TMEM lane quadrant)
//   warps 4-7   : epilogue of row-tile slot 1
//   warp 8      : weight producer  (one lane: waits ring_empty, issues cp.async.bulk)
//   warp 9      : MMA issuer       (one lane: tcgen05.mma / tcgen05.commit); owns the TMEM allocation
//
// Each CTA works on TWO 128-row tiles at a time ("slots"). The MMA issuer alternates
// (slot0, job j), (slot1, job j), (slot0, job j+1), ... so that while one slot's epilogue
// (TMEM -> registers -> +bias, ReLU, cast -> swizzled smem = next layer's A operand) runs on the
// CUDA cores, the other slot's layer runs on the tensor core: a layer-granular ping-pong.
//
// Jobs of one tile (K chunks of 64, padded columns are zero in the packed weights):
//   J0  dense_0   A = enc_xyz(63+1)                       N=256
//   J1-4 dense_1..4  A = act(256)                          N=256
//   J5  dense_5   A = act(256) | enc_xyz(64)               N=256   (skip concat = one more K chunk)
//   J6-7 dense_6,7                                          N=256   (+ sigma head in J7's epilogue, fp32)
//   J8  dense_8   linear                                    N=256
//   J9  dense_9   A = bott(256) | enc_dir(27+5)             N=128
//   J10 rgb       A = act(128)                              N=16 (3 used), sigmoid in the epilogue
//
// Shared memory (bytes): 2 x 64 KB activations, 2 x 16 KB encodings, 4 x 16 KB weight ring, barriers.
#include "common.cuh"
#include "mlp.cuh"
#include "tc_common.cuh"
#include "tc_layout.cuh"

#include <stdlib.h>

namespace nb {

// ---------------------------------------------------------------------------------------------
// Packed weight image: a stream of chunks in consumption order. A chunk is [rows x 64 K] 16-bit,
// K-major, 128 bytes per row, 16-byte units XOR-swizzled by (row & 7) (UMMA SWIZZLE_128B).
struct Chunk {
    uint32_t gofs;     // byte offset inside the packed image
    uint16_t rows;     // N rows in this chunk (128 or 16)
    uint8_t layer;     // Layer enum
    uint8_t job;       // 0..10
    uint8_t nh;        // which 128-wide half of N
    uint8_t kc;        // K chunk index within the layer
    uint8_t asrc;      // 0..3 = activation chunk, 4 = encoding buffer
    uint8_t ksteps;    // UMMA K=16 steps in this chunk (4, or 2 for enc_dir)
    uint8_t first;     // first K chunk of this (job, nh): accumulate = 0
    uint8_t last;      // last chunk of the job: commit acc_full
};

constexpr int kNumJobs = 11;
constexpr int kMaxChunks = 80;

constexpr uint32_t kBiasTileBytes = 2048;   // one CTA's share of a layer's bias as a tensor-core B operand (see pack_bias_tiles_kernel)
struct ChunkTable {
    Chunk c[kMaxChunks];
    int n;
    int job_begin[kNumJobs + 1];
    uint32_t bias_ofs;       // byte offset of the bias tiles [job][cta rank][kBiasTileBytes] behind the weight chunks
    uint32_t bytes;
};

static ChunkTable build_chunk_table() {
    ChunkTable t{};
    uint32_t ofs = 0;
    int n = 0;
    auto add = [&](int job, int layer, int nh, int kc, int rows, int asrc, int ksteps, bool first) {
        Chunk& c = t.c[n++];
        c.gofs = ofs; c.rows = (uint16_t)rows; c.layer = (uint8_t)layer; c.job = (uint8_t)job; c.nh = (uint8_t)nh;
        c.kc = (uint8_t)kc; c.asrc = (uint8_t)asrc; c.ksteps = (uint8_t)ksteps; c.first = first; c.last = 0;
        ofs += (uint32_t)rows * 128u;
    };
    const int job_layer[kNumJobs] = {L0, L1, L2, L3, L4, L5, L6, L7, L8, L9, LRGB};
    for (int j = 0; j < kNumJobs; ++j) {
        t.job_begin[j] = n;
        int l = job_layer[j];
        if (j == 0) {
            for (int nh = 0; nh < 2; ++nh) add(j, l, nh, 0, 128, 4, 4, true);
        } else if (j == 5) {
            for (int nh = 0; nh < 2; ++nh)
                for (int kc = 0; kc < 5; ++kc) add(j, l, nh, kc, 128, kc < 4 ? kc : 4, 4, kc == 0);
        } else if (j == 9) {
            for (int kc = 0; kc < 5; ++kc) add(j, l, 0, kc, 128, kc < 4 ? kc : 4, kc < 4 ? 4 : 2, kc == 0);
        } else if (j == 10) {
            for (int kc = 0; kc < 2; ++kc) add(j, l, 0, kc, 16, kc, 4, kc == 0);
        } else {
            for (int nh = 0; nh < 2; ++nh)
                for (int kc = 0; kc < 4; ++kc) add(j, l, nh, kc, 128, kc, 4, kc == 0);
        }
        t.c[n - 1].last = 1;
    }
    t.job_begin[kNumJobs] = n;
    t.n = n;
    t.bias_ofs = ofs;
    t.bytes = ofs + kNumJobs * 2 * kBiasTileBytes;
    return t;
}

static const ChunkTable& chunk_table() {
    static ChunkTable t = build_chunk_table();
    return t;
}

__constant__ Chunk c_chunks[kMaxChunks];
__constant__ int c_job_begin[kNumJobs + 1];

// fp32 side parameters per model (biases of the 10 tensor-core layers + the sigma head)
struct HeadOffsets {
    // float offsets inside head_params
    __host__ __device__ static constexpr int bias(int job) {   // jobs 0..8 -> 256 each, 9 -> 128, 10 -> 16 (3 used)
        return job <= 9 ? job * 256 : 9 * 256 + 128;
    }
    static constexpr int wsigma = 9 * 256 + 128 + 16;   // 256 floats
    static constexpr int bsigma = wsigma + 256;         // 1 float (+3 pad)
    static constexpr int total = bsigma + 4;
};

template <typename T> __device__ __forceinline__ T to16(float v);
template <> __device__ __forceinline__ __nv_bfloat16 to16<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half to16<__half>(float v) { return __float2half_rn(v); }

// One thread per 16-byte unit of the packed image.
template <typename T, bool kLo>
__device__ __forceinline__ void pack_weights_unit(const float* __restrict__ P /* one model */, uint8_t* __restrict__ img, int nchunks,
                                                  int64_t u) {
    // locate chunk by linear scan over the table (<= 80 entries)
    int ci = -1;
    uint32_t byte = (uint32_t)(u * 16);
    for (int i = 0; i < nchunks; ++i) {
        uint32_t b = c_chunks[i].gofs, e = b + c_chunks[i].rows * 128u;
        if (byte >= b && byte < e) { ci = i; break; }
    }
    if (ci < 0) return;
    const Chunk ch = c_chunks[ci];
    uint32_t local = byte - ch.gofs;
    int row = local >> 7;
    int phys_unit = (local >> 4) & 7;
    int unit = phys_unit ^ (row & 7);
    const LayerDim dim = layer_dim(ch.layer);
    const float* W = P + kernel_offset(ch.layer);
    const float* bias = P + bias_offset(ch.layer);
    (void)bias;
    int n = ch.nh * 128 + row;
    T vals[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        int kl = unit * 8 + e;         // 0..63 inside the chunk
        int k;                         // row of the Keras kernel [in,out], or -1 for zero padding
        if (ch.layer == L0) k = kl < 63 ? kl : -1;
        else if (ch.layer == L5) k = ch.kc < 4 ? ch.kc * 64 + kl : (kl < 63 ? 256 + kl : -1);
        else if (ch.layer == L9) k = ch.kc < 4 ? ch.kc * 64 + kl : (kl < 27 ? 256 + kl : -1);
        else k = ch.kc * 64 + kl;
        float v = 0.f;
        if (k >= 0 && k < dim.fan_in && n < dim.fan_out) v = W[(int64_t)k * dim.fan_out + n];
        vals[e] = to16<T>(v);
        if (kLo) vals[e] = to16<T>(v - (float)vals[e]);      // lo image of the split launch: what the 16-bit rounding dropped
    }
    *reinterpret_cast<uint4*>(img + byte) = *reinterpret_cast<uint4*>(vals);
}

// The layer bias as a tensor-core operand (pair kernel): the first MMA of every layer multiplies a "ones" A operand
// (columns k = 0, 1 are 1.0, the rest 0; see the kernel) with this B tile and starts the accumulator at the bias, so
// the epilogue neither loads nor adds it (the broadcast shared-memory loads of the bias were ~1000 of its ~1800
// cycles per layer: as many bytes into registers as the accumulator itself; tools/epi_probe.cu). K-major, no swizzle:
// N-row i of the CTA's share is the 16 bytes [hi, lo, 0, 0, 0, 0, 0, 0] with hi + lo = bias to ~2^-17 relative.
template <typename T>
__device__ __forceinline__ void pack_bias_tile_row(const float* __restrict__ P /* one model */, uint8_t* __restrict__ tiles, int i) {
    if (i >= kNumJobs * 2 * 128) return;                          // i = (job, rank, row)
    const int job = i / 256, rank = (i >> 7) & 1, row = i & 127;
    const int job_layer[kNumJobs] = {L0, L1, L2, L3, L4, L5, L6, L7, L8, L9, LRGB};
    const int l = job_layer[job];
    const int N = job < 9 ? 256 : (job == 9 ? 128 : 16);
    const int rows = N / 2;                                        // N-rows this CTA supplies
    T v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = to16<T>(0.f);
    const int n = rank * rows + row;
    if (row < rows && n < layer_dim(l).fan_out) {
        const float b = P[bias_offset(l) + n];
        v[0] = to16<T>(b);
        v[1] = to16<T>(b - (float)v[0]);
    }
    *reinterpret_cast<uint4*>(tiles + (size_t)(job * 2 + rank) * kBiasTileBytes + row * 16) = *reinterpret_cast<uint4*>(v);
}

__device__ __forceinline__ void pack_heads_elem(const float* __restrict__ P, float* __restrict__ hp, int i) {
    if (i >= HeadOffsets::total) return;
    const int job_layer[kNumJobs] = {L0, L1, L2, L3, L4, L5, L6, L7, L8, L9, LRGB};
    float v = 0.f;
    if (i < HeadOffsets::wsigma) {
        int job = i < 9 * 256 ? i / 256 : (i < 9 * 256 + 128 ? 9 : 10);
        int n = i - HeadOffsets::bias(job);
        int l = job_layer[job];
        if (n < layer_dim(l).fan_out) v = P[bias_offset(l) + n];
    } else if (i < HeadOffsets::bsigma) {
        v = P[kernel_offset(LSIGMA) + (i - HeadOffsets::wsigma)];
    } else if (i == HeadOffsets::bsigma) {
        v = P[bias_offset(LSIGMA)];
    }
    hp[i] = v;
}

// ONE launch packs everything the forward kernels of one precision read, for both models (blockIdx.y): the operand image
// (hi), optionally the lo image of the split launch, the bias tiles and the fp32 head block -- a training step repacks
// after every Adam update, and at 512 rays per GPU ten separate 3-10 us launches were 5 % of a data-parallel step.
struct PackParams {
    const float* P;              // flat parameters, coarse then fine
    uint8_t* img[2];             // operand image per model (bias tiles behind the chunks)
    uint8_t* img_lo[2];          // lo image per model, or NULL
    float* heads[2];
    uint32_t bias_ofs;
    int nchunks;
    int blocks_w, blocks_b, blocks_h;      // 256-thread blocks per section; sections: hi [, lo], bias tiles, heads
};
template <typename T>
__global__ void pack_images_kernel(const PackParams pp) {
    const int m = blockIdx.y;
    const float* P = pp.P + (int64_t)m * kParamsPerModel;
    int b = blockIdx.x;
    if (b < pp.blocks_w) { pack_weights_unit<T, false>(P, pp.img[m], pp.nchunks, (int64_t)b * blockDim.x + threadIdx.x); return; }
    b -= pp.blocks_w;
    if (pp.img_lo[m]) {
        if (b < pp.blocks_w) { pack_weights_unit<T, true>(P, pp.img_lo[m], pp.nchunks, (int64_t)b * blockDim.x + threadIdx.x); return; }
        b -= pp.blocks_w;
    }
    if (b < pp.blocks_b) { pack_bias_tile_row<T>(P, pp.img[m] + pp.bias_ofs, b * blockDim.x + threadIdx.x); return; }
    b -= pp.blocks_b;
    pack_heads_elem(P, pp.heads[m], b * blockDim.x + threadIdx.x);
}

// ---------------------------------------------------------------------------------------------
constexpr int kActBytes = 4 * 16384;    // 128 rows x 256 K x 2 B, four K chunks of [128 x 64]
constexpr int kEncBytes = 16384;        // 128 rows x 64 K x 2 B
constexpr int kStageBytes = 16384;      // [128 N x 64 K]
constexpr int kStages = 4;
constexpr int kSmemAct = 0;
constexpr int kSmemEnc = kSmemAct + 2 * kActBytes;
constexpr int kSmemRing = kSmemEnc + 2 * kEncBytes;
constexpr int kSmemBar = kSmemRing + kStages * kStageBytes;
constexpr int kSmemBias = kSmemBar + 256;             // single-CTA kernel: 2 slots x 256 fp32 bias staging
constexpr int kSmemOnes = kSmemBias;                  // pair kernel: 256 B "ones" A operand + one 2 KB bias tile instead
constexpr int kSmemBiasTile = kSmemOnes + 256;
constexpr int kSmemTotal = kSmemBiasTile + 2048;
static_assert(kSmemTotal <= 232448, "exceeds the 227 KB dynamic shared memory limit");
constexpr int kThreads = 320;
constexpr int kProducerWarp = 8, kMmaWarp = 9;   // highest warp ids: the per-SMSP arbiter favours high warp ids,
                                                 // and the single MMA-issuing thread is the critical path

struct TcParams {
    const uint8_t* wimg;     // packed weights of this model/precision
    const uint8_t* wimg_lo;  // split instantiation: the image of W - fl16(W) (same layout), else NULL
    uint32_t bias_ofs;       // byte offset of the bias tiles inside the image
    int last_only;           // split instantiation: row r of the launch is the LAST sample of ray r (row r*S + S-1)
    unsigned long long* dbg; // optional cycle counters (debug instantiation, nerfb200_set_option), else NULL
    int dbg_mode;            // debug instantiation: 3/4/5 = drop the stash stores / the bitmask stores / both
    const float* heads;      // HeadOffsets block
    const float* ro; const float* rd; const float* t;
    float* rgb; float* sigma;
    uint8_t* stash;          // training: per-tile activation stash (tc_layout.cuh), else NULL
    int64_t R;               // rows
    int S;
    int num_tiles;
};

struct RowCtx {
    int64_t grow;
    bool valid;
    float dir[3];
};

// sin/cos of the L octaves of one coordinate for the tensor-core operands. The reference's arguments
// fl32(x * fl32(2^l pi)) (core/model.py:318-324) are exactly 2^l * a0 with a0 = fl32(x * fl32(pi)) -- scaling by
// a power of two commutes with rounding -- so an accurate sincosf every 5th octave plus the double-angle
// identities in between evaluate the SAME arguments with <= ~3e-6 absolute error (2^4 fp32 ulps), three
// orders of magnitude below the 16-bit rounding the operand gets next (bf16 ulp 3.9e-3, fp16 4.9e-4).
// The fp32 check path and the standalone posenc kernel keep one sincosf per octave.
template <int L>
__device__ __forceinline__ void sincos_octaves(float x, float* e) {
    const float a0 = __fmul_rn(x, 3.14159274101257324f);
    float sn = 0.f, cs = 1.f;
#pragma unroll
    for (int l = 0; l < L; ++l) {
        if (l % 5 == 0) {
            sincosf(__fmul_rn(a0, (float)(1 << l)), &sn, &cs);
        } else {
            const float s2 = __fmul_rn(__fmul_rn(2.f, sn), cs);
            cs = fmaf(__fmul_rn(-2.f, sn), sn, 1.f);
            sn = s2;
        }
        e[2 * l] = sn;
        e[2 * l + 1] = cs;
    }
}

// Loads the ray of this thread's row, forms xyz = o + t*d (utils/ray_utils.py:251) and writes the
// L=10 positional encoding (core/model.py:305-332) into the 64-column encoding buffer (col 63 = 0).
template <bool kHalf, bool kSplit = false>
__device__ __forceinline__ void prep_tile(const TcParams& p, int tile, int row, uint8_t* enc, RowCtx& rc, uint8_t* enc_lo = nullptr) {
    rc.grow = (int64_t)tile * kTileRows + row;
    rc.valid = rc.grow < p.R;
    int64_t lrow = rc.valid ? rc.grow : p.R - 1;
    int64_t ray;
    if (kSplit) {           // the launch's row r is the last sample of ray r
        ray = lrow;
        lrow = lrow * p.S + (p.S - 1);
        rc.grow = lrow;
    } else {
        ray = (p.R <= 0x7fffffffLL) ? (int64_t)((uint32_t)lrow / (uint32_t)p.S) : lrow / p.S;
    }
    const float tv = __ldg(p.t + lrow);
    float xyz[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        rc.dir[d] = __ldg(p.rd + 3 * ray + d);
        xyz[d] = __fadd_rn(__ldg(p.ro + 3 * ray + d), __fmul_rn(tv, rc.dir[d]));
    }
    float e[64];
    e[0] = xyz[0]; e[1] = xyz[1]; e[2] = xyz[2];
#pragma unroll
    for (int d = 0; d < 3; ++d) sincos_octaves<10>(xyz[d], e + 3 + d * 20);
    e[63] = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        uint4 v;
        v.x = pack2<kHalf>(e[8 * u + 0], e[8 * u + 1]);
        v.y = pack2<kHalf>(e[8 * u + 2], e[8 * u + 3]);
        v.z = pack2<kHalf>(e[8 * u + 4], e[8 * u + 5]);
        v.w = pack2<kHalf>(e[8 * u + 6], e[8 * u + 7]);
        *reinterpret_cast<uint4*>(enc + swz(row, u)) = v;
        if (kSplit) {       // the part of the encoding that the 16-bit rounding dropped
            uint4 w;
            w.x = pack2<kHalf>(e[8 * u + 0] - unpack_lo<kHalf>(v.x), e[8 * u + 1] - unpack_hi<kHalf>(v.x));
            w.y = pack2<kHalf>(e[8 * u + 2] - unpack_lo<kHalf>(v.y), e[8 * u + 3] - unpack_hi<kHalf>(v.y));
            w.z = pack2<kHalf>(e[8 * u + 4] - unpack_lo<kHalf>(v.z), e[8 * u + 5] - unpack_hi<kHalf>(v.z));
            w.w = pack2<kHalf>(e[8 * u + 6] - unpack_lo<kHalf>(v.w), e[8 * u + 7] - unpack_hi<kHalf>(v.w));
            *reinterpret_cast<uint4*>(enc_lo + swz(row, u)) = w;
        }
    }
    fence_proxy_async();
}

// enc_dir (L=4) into the encoding buffer, columns 27..63 = 0.
template <bool kHalf>
__device__ __forceinline__ void write_enc_dir(const float (&dir)[3], uint8_t* enc, int row) {
    float e[32];
    e[0] = dir[0]; e[1] = dir[1]; e[2] = dir[2];
#pragma unroll
    for (int d = 0; d < 3; ++d) sincos_octaves<4>(dir[d], e + 3 + d * 8);
#pragma unroll
    for (int i = 27; i < 32; ++i) e[i] = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        uint4 v4 = make_uint4(0u, 0u, 0u, 0u);
        if (u < 4) {
            v4.x = pack2<kHalf>(e[8 * u + 0], e[8 * u + 1]);
            v4.y = pack2<kHalf>(e[8 * u + 2], e[8 * u + 3]);
            v4.z = pack2<kHalf>(e[8 * u + 4], e[8 * u + 5]);
            v4.w = pack2<kHalf>(e[8 * u + 6], e[8 * u + 7]);
        }
        *reinterpret_cast<uint4*>(enc + swz(row, u)) = v4;
    }
}


// Inference epilogue with packed arithmetic: add.f32x2 for the bias and ONE conversion per pair with the
// ReLU folded into it (cvt.rn.relu; rounding is monotonic and sign preserving, so it equals cvt(max(x, 0))).
template <bool kHalf, int NG, bool kRelu, bool kSigma = false, bool kBias = true>
__device__ __forceinline__ void epilogue_cols_packed(uint32_t tmem_row, const float* s_bias, uint8_t* act, int row,
                                                     uint32_t* mask_row = nullptr, const float4* ws4 = nullptr,
                                                     float* sig_acc = nullptr) {
    uint32_t r[2][32];
    tmem_ld32(tmem_row, r[0]);
    float sg[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t mq[4] = {0u, 0u, 0u, 0u};   // training: four mask words at a time, ONE 16-byte store per 128 columns
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        uint32_t (&rr)[32] = r[g & 1];
        tmem_ld_wait(rr);
        if (g + 1 < NG) tmem_ld32(tmem_row + (uint32_t)(32 * (g + 1)), r[(g + 1) & 1]);
        const float4* b4 = reinterpret_cast<const float4*>(s_bias + 32 * g);
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float2 s0 = make_float2(__uint_as_float(rr[4 * i + 0]), __uint_as_float(rr[4 * i + 1]));
            float2 s1 = make_float2(__uint_as_float(rr[4 * i + 2]), __uint_as_float(rr[4 * i + 3]));
            if (kBias) {     // (the pair kernel's accumulators already contain the bias)
                const float4 bb = b4[i];
                s0 = __fadd2_rn(s0, make_float2(bb.x, bb.y));
                s1 = __fadd2_rn(s1, make_float2(bb.z, bb.w));
            }
            if (kSigma) {   // sigma head on the fp32 activations (core/model.py:375); four independent partial sums:
                            // one 256-long dependent FMA chain was ~5 k cycles on the slot's critical path
                const float4 w = __ldg(ws4 + 8 * g + i);
                sg[0] = fmaf(fmaxf(s0.x, 0.f), w.x, sg[0]);
                sg[1] = fmaf(fmaxf(s0.y, 0.f), w.y, sg[1]);
                sg[2] = fmaf(fmaxf(s1.x, 0.f), w.z, sg[2]);
                sg[3] = fmaf(fmaxf(s1.y, 0.f), w.w, sg[3]);
            }
            if (kRelu) { o[2 * i] = pack2_relu<kHalf>(s0.x, s0.y); o[2 * i + 1] = pack2_relu<kHalf>(s1.x, s1.y); }
            else { o[2 * i] = pack2<kHalf>(s0.x, s0.y); o[2 * i + 1] = pack2<kHalf>(s1.x, s1.y); }
        }
        if (kRelu && mask_row) {   // training: ReLU bitmask from the packed outputs (layout: mask_bit(), tc_layout.cuh)
            uint32_t m = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) m |= pos_mask2<kHalf>(o[k]) & (0x00010001u << k);
            mq[g & 3] = m;
            // (eight separate 4-byte stores per row and layer, 32 bytes apart across the lanes, cost the training
            // forward 0.12 ms of 1.1: tools/fwd_train_bench.py, NERFB200_OPT_DEBUG mode 4)
            if ((g & 3) == 3) *reinterpret_cast<uint4*>(mask_row + (g - 3)) = make_uint4(mq[0], mq[1], mq[2], mq[3]);
        }
        uint8_t* chunk = act + (g >> 1) * 16384;
        const int u0 = (g & 1) * 4;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            *reinterpret_cast<uint4*>(chunk + swz(row, u0 + u)) = make_uint4(o[4 * u], o[4 * u + 1], o[4 * u + 2], o[4 * u + 3]);
    }
    if (kSigma) *sig_acc += (sg[0] + sg[1]) + (sg[2] + sg[3]);
}

// Split instantiation (last-sample rows, see mlp_tc_forward_pair_kernel): the post-ReLU activation leaves as TWO
// 16-bit images, hi = fl16(v) and lo = fl16(v - hi), so that the next layer's three MMAs hi.Whi + lo.Whi + hi.Wlo
// carry ~16 (bf16) / ~22 (fp16) significand bits instead of 8 / 11.
template <bool kHalf, bool kSigma>
__device__ __forceinline__ void epilogue_cols_split(uint32_t tmem_row, uint8_t* act_hi, uint8_t* act_lo, int row,
                                                    const float4* ws4, float* sig_acc) {
    uint32_t r[2][32];
    tmem_ld32(tmem_row, r[0]);
    float sg[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        uint32_t (&rr)[32] = r[g & 1];
        tmem_ld_wait(rr);
        if (g + 1 < 8) tmem_ld32(tmem_row + (uint32_t)(32 * (g + 1)), r[(g + 1) & 1]);
        uint32_t oh[16], ol[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float v0 = fmaxf(__uint_as_float(rr[4 * i + 0]), 0.f), v1 = fmaxf(__uint_as_float(rr[4 * i + 1]), 0.f);
            const float v2 = fmaxf(__uint_as_float(rr[4 * i + 2]), 0.f), v3 = fmaxf(__uint_as_float(rr[4 * i + 3]), 0.f);
            if (kSigma) {   // sigma head on the fp32 activations (core/model.py:375)
                const float4 w = __ldg(ws4 + 8 * g + i);
                sg[0] = fmaf(v0, w.x, sg[0]); sg[1] = fmaf(v1, w.y, sg[1]);
                sg[2] = fmaf(v2, w.z, sg[2]); sg[3] = fmaf(v3, w.w, sg[3]);
            }
            oh[2 * i] = pack2<kHalf>(v0, v1);
            oh[2 * i + 1] = pack2<kHalf>(v2, v3);
            ol[2 * i] = pack2<kHalf>(v0 - unpack_lo<kHalf>(oh[2 * i]), v1 - unpack_hi<kHalf>(oh[2 * i]));
            ol[2 * i + 1] = pack2<kHalf>(v2 - unpack_lo<kHalf>(oh[2 * i + 1]), v3 - unpack_hi<kHalf>(oh[2 * i + 1]));
        }
        const int cofs = (g >> 1) * 16384, u0 = (g & 1) * 4;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            *reinterpret_cast<uint4*>(act_hi + cofs + swz(row, u0 + u)) = make_uint4(oh[4 * u], oh[4 * u + 1], oh[4 * u + 2], oh[4 * u + 3]);
            *reinterpret_cast<uint4*>(act_lo + cofs + swz(row, u0 + u)) = make_uint4(ol[4 * u], ol[4 * u + 1], ol[4 * u + 2], ol[4 * u + 3]);
        }
    }
    if (kSigma) *sig_acc += (sg[0] + sg[1]) + (sg[2] + sg[3]);
}

#define NB_T0() long long _t0 = dbg_on ? clock64() : 0
#define NB_T1(slot) do { if (dbg_on) dbg_acc##slot += (unsigned long long)(clock64() - _t0); } while (0)


// =============================================================================================
// 2-CTA variant (tcgen05 cta_group::2): the two CTAs of a cluster (an SM pair) run ONE M=256 MMA per
// instruction -- each CTA contributes its own 128-row tile as A and HALF of the layer's weights as B --
// so per SM the weight stream, the operand reads from shared memory and the number of MMA instructions
// all halve, and every instruction carries 128 cycles of tensor work (above the ~78-cycle issue floor).
// Roles per CTA are as in the single-CTA kernel; only the leader's warp 9 issues MMAs, the peer's warp 9
// relays "my half of the weight stage has landed" to the leader's ring barrier. tcgen05.commit multicasts
// completion to both CTAs' barriers.
//
// kSplit instantiation -- the last sample of every ray at fp32-grade accuracy. The reference sets delta_last = 1e10
// (utils/ray_utils.py:459-468), so alpha_last jumps from 0 to 1 when the last sample's ReLU'd sigma leaves zero: a
// 16-bit rounding error that flips that ONE sign moves the pixel by T_last * (rgb_last - background). A second,
// small launch therefore recomputes sigma of row ray*S + S-1 of every ray (1 of 64 / 192 rows) through dense_0..7
// and the sigma head with error-compensated operands: activations and weights are split into hi + lo 16-bit
// parts (slot 0's buffers hold hi, slot 1's hold lo, of ONE tile per CTA) and every K chunk issues hi.Whi,
// lo.Whi, hi.Wlo into the same accumulator. Only the sigma of those rows is overwritten.
template <bool kHalf, bool kTrain, bool kDbg, bool kSplit = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) mlp_tc_forward_pair_kernel(const TcParams p) {
    static_assert(!(kSplit && (kTrain || kDbg)), "the split instantiation is inference-shaped (it patches sigma only)");
    constexpr int kJobs = kSplit ? 8 : kNumJobs;        // split: dense_0..7 (+ sigma head in the last epilogue)
    constexpr int kTilesPerCluster = kSplit ? 2 : 4;    // split: one tile per CTA (slot 1's buffers hold the lo parts)
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    // cycle counters exist only in the kDbg instantiation: even predicated-off code in the issuer loop costs throughput
    const bool dbg_on = kDbg && p.dbg != nullptr && blockIdx.x < 2;
    unsigned long long dbg_acc0 = 0, dbg_acc1 = 0, dbg_acc2 = 0, dbg_acc3 = 0;
    const long long t_kernel0 = clock64();
    const uint32_t sbase = smem_u32(smem);
    const uint32_t sbar = sbase + kSmemBar;
    auto ring_full = [&](int s) { return sbar + 8 * s; };
    auto ring_empty = [&](int s) { return sbar + 8 * (kStages + s); };
    auto act_ready = [&](int t) { return sbar + 8 * (2 * kStages + t); };
    auto acc_full = [&](int t) { return sbar + 8 * (2 * kStages + 2 + t); };
    const uint32_t bias_full = sbar + 8 * (2 * kStages + 4), bias_empty = sbar + 8 * (2 * kStages + 5);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kSmemBar + 8 * (2 * kStages + 6));

    if (threadIdx.x == 0) {
        // leader's ring_full: own expect_tx arrive + the peer's relay; peer's ring_full: own arrive only
        for (int s = 0; s < kStages; ++s) { mbar_init(ring_full(s), rank == 0 ? 2 : 1); mbar_init(ring_empty(s), 1); }
        // act_ready (used in the leader): one elected arrive per CTA; acc_full: one multicast commit
        for (int t = 0; t < 2; ++t) { mbar_init(act_ready(t), 2); mbar_init(acc_full(t), 1); }
        mbar_init(bias_full, rank == 0 ? 2 : 1);     // like ring_full: own expect_tx arrive (+ the peer's relay in the leader)
        mbar_init(bias_empty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 16) {
        // "ones" A operand of the bias MMA step: core matrix 0 = 8 rows x [1, 1, 0, 0, 0, 0, 0, 0], core matrix 1 = zeros;
        // its descriptor has SBO = 0, so all sixteen 8-row groups of the 128-row tile read these same 256 bytes
        const uint32_t one2 = pack2<kHalf>(1.f, 1.f);
        reinterpret_cast<uint4*>(smem + kSmemOnes)[threadIdx.x] = threadIdx.x < 8 ? make_uint4(one2, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async();
    }
    cluster_sync_all();
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int num_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
    const int quads = (p.num_tiles + kTilesPerCluster - 1) / kTilesPerCluster;   // a cluster works on 4 tiles at a time: 2 slots x 2 CTAs
    constexpr int fmt = kHalf ? 0 : 1;
    auto slots_of = [&](int qd) { return kSplit ? 1 : ((qd * 4 + 2 < p.num_tiles) ? 2 : 1); };   // slots with a tile in at least one CTA

    if (warp == kProducerWarp) {
        // ===================== weight producer: this CTA's half of every chunk =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, bphase = 0;
            for (int qd = cluster_id; qd < quads; qd += num_clusters) {
                const int nslots = slots_of(qd);
                for (int j = 0; j < kJobs; ++j) {
                    {   // this CTA's share of the layer's bias tile (single buffer, released by the last slot's bias MMA)
                        const uint32_t bbytes = j < 9 ? 2048u : (j == 9 ? 1024u : 128u);
                        mbar_wait(bias_empty, bphase ^ 1);
                        mbar_expect_tx(bias_full, bbytes);
                        bulk_g2s(sbase + kSmemBiasTile, p.wimg + p.bias_ofs + (uint32_t)(j * 2 + (int)rank) * kBiasTileBytes, bbytes, bias_full);
                        bphase ^= 1;
                    }
                    const int cb = c_job_begin[j];
                    const int KC = (j == 0) ? 1 : (j == 5 || j == 9) ? 5 : (j == 10) ? 2 : 4;
                    // a layer whose weights fit the ring (KC <= 4) is loaded ONCE and used by both slots
                    const int loads = (KC <= kStages) ? 1 : nslots;
                    for (int rep = 0; rep < loads; ++rep)
                        for (int kc = 0; kc < KC; ++kc) {
                            // N = 256 layers: chunk (nh = rank, kc); dense_9 / rgb: my half of the rows of chunk kc
                            uint32_t gofs, bytes;
                            if (j < 9) { gofs = c_chunks[cb + (int)rank * KC + kc].gofs; bytes = 16384; }
                            else if (j == 9) { gofs = c_chunks[cb + kc].gofs + rank * 8192u; bytes = 8192; }
                            else { gofs = c_chunks[cb + kc].gofs + rank * 1024u; bytes = 1024; }
                            mbar_wait(ring_empty(stage), phase ^ 1);
                            mbar_expect_tx(ring_full(stage), bytes);
                            bulk_g2s(sbase + kSmemRing + stage * kStageBytes, p.wimg + gofs, bytes, ring_full(stage));
                            if (++stage == kStages) { stage = 0; phase ^= 1; }
                            if (kSplit) {      // the same chunk of the lo image goes into the next stage
                                mbar_wait(ring_empty(stage), phase ^ 1);
                                mbar_expect_tx(ring_full(stage), bytes);
                                bulk_g2s(sbase + kSmemRing + stage * kStageBytes, p.wimg_lo + gofs, bytes, ring_full(stage));
                                if (++stage == kStages) { stage = 0; phase ^= 1; }
                            }
                        }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        if (lane == 0 && rank == 1) {
            // ===================== peer: relay "stage landed" to the leader =====================
            uint32_t stage = 0, phase = 0, bphase = 0;
            const uint32_t bias_full_leader = mapa(bias_full, 0);
            for (int qd = cluster_id; qd < quads; qd += num_clusters) {
                const int nslots = slots_of(qd);
                for (int j = 0; j < kJobs; ++j) {
                    mbar_wait(bias_full, bphase);
                    mbar_arrive_cluster(bias_full_leader);
                    bphase ^= 1;
                    const int KC = (j == 0) ? 1 : (j == 5 || j == 9) ? 5 : (j == 10) ? 2 : 4;
                    const int loads = ((KC <= kStages) ? 1 : nslots) * KC * (kSplit ? 2 : 1);
                    for (int c = 0; c < loads; ++c) {
                        mbar_wait(ring_full(stage), phase);
                        mbar_arrive_cluster(mapa(ring_full(stage), 0));
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        } else if (rank == 0) {
            // ===================== leader: MMA issuer for the pair =====================
            // the whole warp walks the loop (converged, so descriptors live in uniform registers and the
            // compiler needs no per-instruction election loop); one elected lane issues the tcgen05 ops
            const uint32_t ring_lo = ((sbase + kSmemRing) >> 4) & 0x3FFFu;
            constexpr uint32_t id256 = umma_idesc_pair(fmt, 256), id128 = umma_idesc_pair(fmt, 128), id16 = umma_idesc_pair(fmt, 16);
            uint32_t stage = 0, phase = 0, act_phase_bits = 0, bphase = 0;
            // no-swizzle K-major descriptors of the bias step: A = the ones atom (LBO 128 B between its two core matrices,
            // SBO 0: every 8-row group aliases it), B = the bias tile (SBO 128 B between 8-row groups, LBO 0: k 8..15 alias
            // k 0..7 and meet A's zeros)
            const uint64_t ones_desc = (uint64_t)(((sbase + kSmemOnes) >> 4) & 0x3FFFu) | ((uint64_t)(128u >> 4) << 16) | (1ull << 46);
            const uint64_t biast_desc = (uint64_t)(((sbase + kSmemBiasTile) >> 4) & 0x3FFFu) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
            for (int qd = cluster_id; qd < quads; qd += num_clusters) {
#pragma unroll 1
                for (int j = 0; j < kJobs; ++j) {
                    const int KC = (j == 0) ? 1 : (j == 5 || j == 9) ? 5 : (j == 10) ? 2 : 4;
                    const int enc_kc = (j == 0) ? 0 : (j == 5 || j == 9) ? 4 : -1;
                    const bool enc_short = (j == 9);
                    const uint32_t idesc = (j < 9) ? id256 : (j == 9) ? id128 : id16;
                    const bool shared_w = KC <= kStages;                     // weights loaded once for both slots
                    const int nslots = slots_of(qd);
                    const uint32_t stage0 = stage, phase0 = phase;
#pragma unroll 1
                    for (int t = 0; t < nslots; ++t) {
                        { NB_T0(); mbar_wait_cluster(act_ready(t), (act_phase_bits >> t) & 1u); NB_T1(0); }
                        act_phase_bits ^= 1u << t;
                        tc_fence_after();
                        const uint32_t act_lo = ((sbase + kSmemAct + t * kActBytes) >> 4) & 0x3FFFu;
                        const uint32_t enc_lo = ((sbase + kSmemEnc + t * kEncBytes) >> 4) & 0x3FFFu;
                        const uint32_t d = tmem_base + (uint32_t)(t * 256);
                        if (shared_w) { stage = stage0; phase = phase0; }     // second slot re-walks the same stages
                        const bool first_user = !shared_w || t == 0, last_user = !shared_w || t == nslots - 1;
                        // the layer starts from its bias: D = ones . bias_tile^T (accumulate off)
                        if (t == 0) { NB_T0(); mbar_wait_cluster(bias_full, bphase); NB_T1(1); tc_fence_after(); bphase ^= 1; }
                        // ONE election per (layer, slot): the elected lane walks the K chunks alone (waits, MMAs, commits);
                        // an election + reconvergence per chunk costs ~200 cycles of issue time (tools/umma_probe.cu)
                        if (elect_one_sync()) {
                            umma_f16_pair(d, ones_desc, biast_desc, idesc, 0u);
                            if (t == nslots - 1) umma_commit_pair(bias_empty);
                            uint32_t st = stage, ph = phase;
#pragma unroll 1
                            for (int kc = 0; kc < KC; ++kc) {
                                if (first_user) { NB_T0(); mbar_wait_cluster(ring_full(st), ph); NB_T1(1); tc_fence_after(); }
                                const bool is_enc = kc == enc_kc;
                                const uint32_t a_lo = is_enc ? enc_lo : act_lo + (uint32_t)(kc * 1024);
                                const uint32_t b_lo = ring_lo + st * (kStageBytes >> 4);
                                umma_f16_pair(d, umma_desc_from_lo(a_lo), umma_desc_from_lo(b_lo), idesc, 1u);
                                umma_f16_pair(d, umma_desc_from_lo(a_lo + 2), umma_desc_from_lo(b_lo + 2), idesc, 1u);
                                if (!(is_enc && enc_short)) {
                                    umma_f16_pair(d, umma_desc_from_lo(a_lo + 4), umma_desc_from_lo(b_lo + 4), idesc, 1u);
                                    umma_f16_pair(d, umma_desc_from_lo(a_lo + 6), umma_desc_from_lo(b_lo + 6), idesc, 1u);
                                }
                                if (kSplit) {
                                    // lo(A) . hi(W): slot 1's buffers hold the lo parts of this tile; then hi(A) . lo(W)
                                    // from the next stage (the split launch never sees the short enc_dir chunk)
                                    const uint32_t l_lo = is_enc ? enc_lo + (uint32_t)(kEncBytes >> 4) : a_lo + (uint32_t)(kActBytes >> 4);
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        umma_f16_pair(d, umma_desc_from_lo(l_lo + 2 * k), umma_desc_from_lo(b_lo + 2 * k), idesc, 1u);
                                    umma_commit_pair(ring_empty(st));
                                    if (++st == kStages) { st = 0; ph ^= 1; }
                                    mbar_wait_cluster(ring_full(st), ph);
                                    tc_fence_after();
                                    const uint32_t b2_lo = ring_lo + st * (kStageBytes >> 4);
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        umma_f16_pair(d, umma_desc_from_lo(a_lo + 2 * k), umma_desc_from_lo(b2_lo + 2 * k), idesc, 1u);
                                }
                                if (last_user) umma_commit_pair(ring_empty(st));
                                if (kc == KC - 1) umma_commit_pair(acc_full(t));
                                if (++st == kStages) { st = 0; ph ^= 1; }
                            }
                        }
                        __syncwarp();
                        // every lane advances the ring position by KC stages (2 KC in the split instantiation)
                        const uint32_t used = (uint32_t)KC * (kSplit ? 2u : 1u);
                        phase ^= ((stage + used) / kStages) & 1u;
                        stage = (stage + used) % kStages;
                    }
                }
            }
        }
    } else if (kSplit) {
        // ===================== split instantiation: warps 0-3 are the epilogue of the CTA's one tile ==========
        if (warp < 4) {
            const int q = warp & 3, row = q * 32 + lane;
            uint8_t* act_hi = smem + kSmemAct;
            uint8_t* act_lo = smem + kSmemAct + kActBytes;
            uint8_t* enc_hi = smem + kSmemEnc;
            uint8_t* enc_lo = smem + kSmemEnc + kEncBytes;
            const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16);
            const uint32_t act_ready_leader = mapa(act_ready(0), 0);
            uint32_t acc_phase = 0;
            const float4* ws4 = reinterpret_cast<const float4*>(p.heads + HeadOffsets::wsigma);
            const float bsig = __ldg(p.heads + HeadOffsets::bsigma);
            RowCtx cur, nxt;
            auto tile_of = [&](int qd) { return qd * 2 + (int)rank; };
            int qd = cluster_id;
            if (qd < quads) prep_tile<kHalf, true>(p, tile_of(qd), row, enc_hi, cur, enc_lo);
            for (; qd < quads; qd += num_clusters) {
                float sig_acc = 0.f;
#pragma unroll 1
                for (int j = 0; j < kJobs; ++j) {
                    named_bar_sync(1, kTileRows);
                    if (row == 0) mbar_arrive_cluster(act_ready_leader);
                    mbar_wait(acc_full(0), acc_phase);
                    acc_phase ^= 1;
                    tc_fence_after();
                    if (j == 7) epilogue_cols_split<kHalf, true>(tmem_row, act_hi, act_lo, row, ws4, &sig_acc);
                    else epilogue_cols_split<kHalf, false>(tmem_row, act_hi, act_lo, row, ws4, &sig_acc);
                    if (j == 7) {
                        const float sg = fmaxf(sig_acc + bsig, 0.f);
                        if (cur.valid) {
                            p.sigma[cur.grow] = sg;
                            if (p.stash)     // training forward: the ReLU gate of the sigma head reads the stashed output
                                reinterpret_cast<float*>(p.stash + (size_t)(cur.grow >> 7) * kStashTileBytes + kStashOutOfs)[3 * 128 + (int)(cur.grow & 127)] = sg;
                        }
                        // dense_5 was the last reader of enc_xyz: the next tile's encoding can go in now
                        const int nq = qd + num_clusters;
                        if (nq < quads) prep_tile<kHalf, true>(p, tile_of(nq), row, enc_hi, nxt, enc_lo);
                    }
                    tc_fence_before();
                    fence_proxy_async();
                }
                cur = nxt;
            }
        }
    } else if (warp < 8) {
        // ===================== epilogue warps (identical in both CTAs) =====================
        const int t = warp >> 2, q = warp & 3, row = q * 32 + lane;
        uint8_t* act = smem + kSmemAct + t * kActBytes;
        uint8_t* enc = smem + kSmemEnc + t * kEncBytes;
        const float* s_bias = nullptr;      // the accumulators already contain the bias (first MMA of every layer)
        const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 256);
        const uint32_t act_ready_leader = mapa(act_ready(t), 0);
        uint32_t acc_phase = 0;
        const float4* ws4 = reinterpret_cast<const float4*>(p.heads + HeadOffsets::wsigma);
        RowCtx cur, nxt;
        uint8_t* pendA_dst = nullptr; uint32_t pendA_bytes = 0;      // training: queued bulk stores (see the single-CTA kernel)
        uint8_t* pendE_dst = nullptr; uint32_t pendE_bytes = 0;
        const uint32_t act_saddr = sbase + kSmemAct + t * kActBytes, enc_saddr = sbase + kSmemEnc + t * kEncBytes;
        // tile of (quad, slot, rank); a tile index past the end is a dummy: rows clamped, nothing stored
        auto tile_of = [&](int qd) { return qd * 4 + t * 2 + (int)rank; };
        int qd = cluster_id;
        if (qd < quads && qd * 4 + t * 2 < p.num_tiles) {
            prep_tile<kHalf>(p, tile_of(qd), row, enc, cur);
            if (kTrain && tile_of(qd) < p.num_tiles) { pendE_dst = p.stash + (size_t)tile_of(qd) * kStashTileBytes + kStashChunkEncXyz * 16384; pendE_bytes = 16384; }
        }
        for (; qd < quads; qd += num_clusters) {
            if (qd * 4 + t * 2 >= p.num_tiles) continue;
            const bool real_tile = tile_of(qd) < p.num_tiles;
            uint8_t* tstash = (kTrain && real_tile) ? p.stash + (size_t)tile_of(qd) * kStashTileBytes : nullptr;
            float sig_acc = 0.f;
            for (int j = 0; j < kJobs; ++j) {
                // job boundary: everything the previous step wrote (encoding / activations) is complete and fenced
                named_bar_sync(1 + t, kTileRows);
                if (row == 0) mbar_arrive_cluster(act_ready_leader);      // this CTA's operand for job j is ready
                bool issued = false;
                // debug instantiation only (NERFB200_OPT_DEBUG modes 3/4/5): drop the stash stores / the bitmask stores / both,
                // to attribute the training forward's slowdown over the inference forward (tools/fwd_train_bench.py)
                const bool no_stash_store = kDbg && (p.dbg_mode == 3 || p.dbg_mode == 5);
                const bool no_mask_store = kDbg && (p.dbg_mode == 4 || p.dbg_mode == 5);
                if (kTrain && row == 0 && !no_stash_store) {
                    if (pendA_bytes) { bulk_s2g(pendA_dst, act_saddr, pendA_bytes); issued = true; }
                    if (pendE_bytes) { bulk_s2g(pendE_dst, enc_saddr, pendE_bytes); issued = true; }
                    if (issued) bulk_commit_group();
                }
                pendA_bytes = 0; pendE_bytes = 0;
                { NB_T0(); mbar_wait(acc_full(t), acc_phase); NB_T1(0); }
                acc_phase ^= 1;
                tc_fence_after();
                if (kTrain) {
                    if (issued) bulk_wait_read_all();
                    named_bar_sync(1 + t, kTileRows);
                }
                long long _te = dbg_on ? clock64() : 0;
                if (j < 10) {
                    if (kTrain) {
                        uint32_t* mrow = nullptr;
                        if (tstash && j != 8 && !no_mask_store) mrow = reinterpret_cast<uint32_t*>(tstash + kStashMaskOfs) + ((j == 9 ? 8 : j) * 128 + row) * 8;
                        if (j == 7) epilogue_cols_packed<kHalf, 8, true, true, false>(tmem_row, s_bias, act, row, mrow, ws4, &sig_acc);
                        else if (j == 8) epilogue_cols_packed<kHalf, 8, false, false, false>(tmem_row, s_bias, act, row);
                        else if (j == 9) epilogue_cols_packed<kHalf, 4, true, false, false>(tmem_row, s_bias, act, row, mrow);
                        else epilogue_cols_packed<kHalf, 8, true, false, false>(tmem_row, s_bias, act, row, mrow);
                        if (tstash) {
                            pendA_dst = tstash + (j < 8 ? stash_chunk_Y(j) : j == 8 ? kStashChunkBott : kStashChunkY9) * 16384;
                            pendA_bytes = j == 9 ? 2 * 16384 : 4 * 16384;
                        }
                    } else {
                        if (j == 7) epilogue_cols_packed<kHalf, 8, true, true, false>(tmem_row, s_bias, act, row, nullptr, ws4, &sig_acc);
                        else if (j == 8) epilogue_cols_packed<kHalf, 8, false, false, false>(tmem_row, s_bias, act, row);
                        else if (j == 9) epilogue_cols_packed<kHalf, 4, true, false, false>(tmem_row, s_bias, act, row);
                        else epilogue_cols_packed<kHalf, 8, true, false, false>(tmem_row, s_bias, act, row);
                    }
                    if (j == 5) {
                        long long _td = dbg_on ? clock64() : 0;
                        write_enc_dir<kHalf>(cur.dir, enc, row);
                        if (dbg_on) dbg_acc2 += (unsigned long long)(clock64() - _td);
                        if (kTrain && tstash) { pendE_dst = tstash + kStashChunkEncDir * 16384; pendE_bytes = 16384; }
                    }
                    if (j == 7) {
                        const float sg = fmaxf(sig_acc + __ldg(p.heads + HeadOffsets::bsigma), 0.f);
                        if (cur.valid) p.sigma[cur.grow] = sg;
                        if (kTrain && tstash) reinterpret_cast<float*>(tstash + kStashOutOfs)[3 * 128 + row] = cur.valid ? sg : 0.f;
                    }
                    tc_fence_before();
                    fence_proxy_async();
                    if (dbg_on) dbg_acc1 += (unsigned long long)(clock64() - _te);
                    if (j == 9) {
                        const int nq = qd + num_clusters;
                        if (nq < quads && nq * 4 + t * 2 < p.num_tiles) {
                            long long _tp = dbg_on ? clock64() : 0;
                            prep_tile<kHalf>(p, tile_of(nq), row, enc, nxt);
                            if (dbg_on) dbg_acc3 += (unsigned long long)(clock64() - _tp);
                            if (kTrain && tile_of(nq) < p.num_tiles) { pendE_dst = p.stash + (size_t)tile_of(nq) * kStashTileBytes + kStashChunkEncXyz * 16384; pendE_bytes = 16384; }
                        }
                    }
                } else {
                    uint32_t r[32];
                    tmem_ld32(tmem_row, r);
                    tmem_ld_wait(r);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float x = __uint_as_float(r[c]);
                        const float y = 1.f / (1.f + expf(-x));
                        if (cur.valid) p.rgb[3 * cur.grow + c] = y;
                        if (kTrain && tstash) reinterpret_cast<float*>(tstash + kStashOutOfs)[3 * row + c] = cur.valid ? y : 0.f;
                    }
                    tc_fence_before();
                }
            }
            cur = nxt;
        }
        if (kTrain) {
            named_bar_sync(1 + t, kTileRows);
            if (row == 0) {
                if (pendA_bytes) bulk_s2g(pendA_dst, act_saddr, pendA_bytes);
                if (pendE_bytes) bulk_s2g(pendE_dst, enc_saddr, pendE_bytes);
                bulk_commit_group();
                bulk_wait_all();
            }
        }
    }
    if (dbg_on && lane == 0 && (warp == kProducerWarp || warp == kMmaWarp || warp == 0 || warp == 4)) {
        int rowi = (warp == kProducerWarp ? 0 : warp == kMmaWarp ? 1 : warp == 0 ? 2 : 3);
        unsigned long long* o = p.dbg + blockIdx.x * 32 + rowi * 8;
        o[0] = dbg_acc0; o[1] = dbg_acc1; o[2] = dbg_acc2; o[3] = dbg_acc3; o[4] = (unsigned long long)(clock64() - t_kernel0);
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}



// ---------------------------------------------------------------------------------------------
// __constant__ memory is per device: the tables are uploaded by every tc_create, on the context's device.
static int upload_table() {
    const ChunkTable& t = chunk_table();
    NB_CUDA(cudaMemcpyToSymbol(c_chunks, t.c, sizeof(Chunk) * kMaxChunks));
    NB_CUDA(cudaMemcpyToSymbol(c_job_begin, t.job_begin, sizeof(int) * (kNumJobs + 1)));
    return 0;
}

int check_device(const nerfb200_ctx* ctx, const char* fn) {
    int dev = -1;
    NB_CUDA(cudaGetDevice(&dev));
    NB_CHECK_ARG(dev == ctx->device, "%s: the context was created on device %d but the current device is %d", fn, ctx->device, dev);
    return 0;
}

template <bool H, bool T, bool D, bool S> static int set_smem_attr() {
    NB_CUDA(cudaFuncSetAttribute(mlp_tc_forward_pair_kernel<H, T, D, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    return 0;
}

int tc_create(nerfb200_ctx* ctx) {
    NB_CUDA(cudaGetDevice(&ctx->device));
    NB_CUDA(cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, ctx->device));
    const ChunkTable& t = chunk_table();
    for (int pz = 0; pz < 2; ++pz)
        for (int m = 0; m < 2; ++m) {
            NB_CUDA(cudaMalloc(&ctx->packed[pz][m], (size_t)t.bytes));
            NB_CUDA(cudaMalloc(&ctx->packed_lo[pz][m], (size_t)t.bias_ofs));
        }
    for (int m = 0; m < 2; ++m) NB_CUDA(cudaMalloc((void**)&ctx->head_params[m], HeadOffsets::total * sizeof(float)));
    int rc = upload_table();
    if (rc) return rc;
    if ((rc = set_smem_attr<false, false, false, false>())) return rc;
    if ((rc = set_smem_attr<true, false, false, false>())) return rc;
    if ((rc = set_smem_attr<false, true, false, false>())) return rc;
    if ((rc = set_smem_attr<true, true, false, false>())) return rc;
    if ((rc = set_smem_attr<false, false, true, false>())) return rc;
    if ((rc = set_smem_attr<true, false, true, false>())) return rc;
    if ((rc = set_smem_attr<false, true, true, false>())) return rc;
    if ((rc = set_smem_attr<true, true, true, false>())) return rc;
    if ((rc = set_smem_attr<false, false, false, true>())) return rc;
    if ((rc = set_smem_attr<true, false, false, true>())) return rc;
    if ((rc = tf32_create(ctx))) return rc;
    return tc_train_create(ctx);
}

void tc_destroy(nerfb200_ctx* ctx) {
    tc_train_destroy(ctx);
    tf32_destroy(ctx);
    for (int pz = 0; pz < 2; ++pz)
        for (int m = 0; m < 2; ++m) {
            if (ctx->packed[pz][m]) cudaFree(ctx->packed[pz][m]);
            if (ctx->packed_lo[pz][m]) cudaFree(ctx->packed_lo[pz][m]);
        }
    for (int m = 0; m < 2; ++m) if (ctx->head_params[m]) cudaFree(ctx->head_params[m]);
}

// Packs the operand images of the precisions in ctx->pack_mask (bit 0 bf16, bit 1 fp16, bit 2 tf32; a training loop
// that uses one precision sets the mask and saves the other images' launches every step): one launch per 16-bit precision
// for both models (pack_images_kernel) + one for the W^T images of backward-data (mlp_tc_train.cu).
int tc_pack_weights(nerfb200_ctx* ctx, const float* flat_params, cudaStream_t st) {
    int rc = check_device(ctx, "pack_weights");
    if (rc) return rc;
    const ChunkTable& t = chunk_table();
    PackParams pp{};
    pp.P = flat_params;
    pp.bias_ofs = t.bias_ofs;
    pp.nchunks = t.n;
    pp.blocks_w = (int)((t.bias_ofs / 16 + 255) / 256);
    pp.blocks_b = (kNumJobs * 256 + 255) / 256;
    pp.blocks_h = (HeadOffsets::total + 255) / 256;
    for (int m = 0; m < 2; ++m) pp.heads[m] = ctx->head_params[m];
    // the split launch of a tf32 render uses the bf16 pair of images, so tf32 implies bf16 here
    const bool want[2] = {(ctx->pack_mask & 1) || ((ctx->pack_mask & 4) && ctx->precise_last), (ctx->pack_mask & 2) != 0};
    for (int pz = 0; pz < 2; ++pz) {
        if (!want[pz]) continue;
        for (int m = 0; m < 2; ++m) {
            pp.img[m] = (uint8_t*)ctx->packed[pz][m];
            pp.img_lo[m] = ctx->precise_last ? (uint8_t*)ctx->packed_lo[pz][m] : nullptr;
        }
        const dim3 grid((unsigned)(pp.blocks_w * (ctx->precise_last ? 2 : 1) + pp.blocks_b + pp.blocks_h), 2);
        if (pz == 0) pack_images_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(pp);
        else pack_images_kernel<__half><<<grid, 256, 0, st>>>(pp);
        NB_LAUNCH_CHECK();
    }
    if (!want[0] && !want[1]) {      // tf32 without the split launch: the fp32 head block is still needed
        for (int m = 0; m < 2; ++m) { pp.img[m] = nullptr; pp.img_lo[m] = nullptr; }
        PackParams ph = pp;
        ph.blocks_w = 0; ph.blocks_b = 0;
        pack_images_kernel<__nv_bfloat16><<<dim3((unsigned)ph.blocks_h, 2), 256, 0, st>>>(ph);
        NB_LAUNCH_CHECK();
    }
    if (ctx->pack_mask & 4) { rc = tf32_pack(ctx, flat_params, st); if (rc) return rc; }
    rc = tc_train_pack(ctx, flat_params, st);
    if (rc) return rc;
    ctx->packed_valid = true;
    ctx->packed_mask = ctx->pack_mask | (want[0] ? 1 : 0);
    ctx->packed_precise = ctx->precise_last;
    return 0;
}

// The second, small launch of a forward: sigma of the last sample of every ray with split operands (see the kernel).
static int launch_split(nerfb200_ctx* ctx, int which, int half, int64_t B, int S, const float* ro, const float* rd, const float* t,
                        float* sigma, void* stash, cudaStream_t st) {
    TcParams p{};
    p.wimg = (const uint8_t*)ctx->packed[half ? 1 : 0][which];
    p.wimg_lo = (const uint8_t*)ctx->packed_lo[half ? 1 : 0][which];
    p.bias_ofs = chunk_table().bias_ofs;
    p.last_only = 1;
    p.heads = ctx->head_params[which];
    p.ro = ro; p.rd = rd; p.t = t; p.rgb = nullptr; p.sigma = sigma; p.R = B; p.S = S;
    p.stash = (uint8_t*)stash;
    p.num_tiles = (int)((B + kTileRows - 1) / kTileRows);
    const int units = (p.num_tiles + 1) / 2;
    const int clusters = units < ctx->num_sms / 2 ? units : ctx->num_sms / 2;
    if (half) mlp_tc_forward_pair_kernel<true, false, false, true><<<2 * clusters, kThreads, kSmemTotal, st>>>(p);
    else mlp_tc_forward_pair_kernel<false, false, false, true><<<2 * clusters, kThreads, kSmemTotal, st>>>(p);
    NB_LAUNCH_CHECK();
    return 0;
}

int tc_precise_last(nerfb200_ctx* ctx, int which, int half, int64_t B, int S, const float* ro, const float* rd, const float* t,
                    float* sigma, void* stash, cudaStream_t st) {
    if (!ctx->precise_last || S < 2 || B == 0) return 0;
    if (!ctx->packed_precise || !(ctx->packed_mask & (half ? 2 : 5))) {
        set_error("mlp_forward: the split images of this precision were not packed (option changed after pack_weights)");
        return NERFB200_ESTATE;
    }
    return launch_split(ctx, which, half, B, S, ro, rd, t, sigma, stash, st);
}

int tc_forward(nerfb200_ctx* ctx, int which, int half, int64_t B, int S, const float* ro, const float* rd, const float* t,
               float* rgb, float* sigma, void* workspace, void* stash, cudaStream_t st) {
    (void)workspace;
    int rc = check_device(ctx, "mlp_forward");
    if (rc) return rc;
    if (!ctx->packed_valid || !(ctx->packed_mask & (half ? 2 : 1))) {
        set_error("mlp_forward: pack_weights has not been called for this precision");
        return NERFB200_ESTATE;
    }
    const int64_t R = B * S;
    if (R == 0) return 0;
    NB_CHECK_ARG((R + kTileRows - 1) / kTileRows < (int64_t)1 << 30, "mlp_forward: too many rows");
    TcParams p{};
    p.wimg = (const uint8_t*)ctx->packed[half ? 1 : 0][which];
    p.bias_ofs = chunk_table().bias_ofs;
    p.heads = ctx->head_params[which];
    p.ro = ro; p.rd = rd; p.t = t; p.rgb = rgb; p.sigma = sigma; p.R = R; p.S = S;
    p.stash = (uint8_t*)stash;
    p.num_tiles = (int)((R + kTileRows - 1) / kTileRows);
    const bool debug = ctx->debug != 0;      // developer cycle counters (nerfb200_set_option NERFB200_OPT_DEBUG)
    p.dbg_mode = ctx->debug;
    if (debug) {
        NB_CUDA(cudaMalloc((void**)&p.dbg, 128 * sizeof(unsigned long long)));
        NB_CUDA(cudaMemsetAsync(p.dbg, 0, 128 * sizeof(unsigned long long), st));
    }
    const int quads = (p.num_tiles + 3) / 4;
    const int clusters = quads < ctx->num_sms / 2 ? quads : ctx->num_sms / 2;
#define NB_LAUNCH_PAIR(H, T)                                                                                   \
    do {                                                                                                       \
        if (debug) mlp_tc_forward_pair_kernel<H, T, true><<<2 * clusters, kThreads, kSmemTotal, st>>>(p);      \
        else mlp_tc_forward_pair_kernel<H, T, false><<<2 * clusters, kThreads, kSmemTotal, st>>>(p);           \
    } while (0)
    if (stash) {
        if (half) NB_LAUNCH_PAIR(true, true); else NB_LAUNCH_PAIR(false, true);
    } else {
        if (half) NB_LAUNCH_PAIR(true, false); else NB_LAUNCH_PAIR(false, false);
    }
#undef NB_LAUNCH_PAIR
    NB_LAUNCH_CHECK();
    if (debug) {
        unsigned long long h[128];
        NB_CUDA(cudaMemcpyAsync(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
        NB_CUDA(cudaStreamSynchronize(st));
        cudaFree(p.dbg);
        const char* names[4] = {"producer", "mma", "epi0", "epi1"};
        fprintf(stderr, "[tc debug] tiles=%d clusters=%d (cluster 0 cycles)\n", p.num_tiles, clusters);
        for (int r = 0; r < 8; ++r)
            fprintf(stderr, "  cta%d %-8s wait0=%llu wait1/epi=%llu encdir=%llu prep=%llu total=%llu\n", r / 4, names[r % 4], h[r * 8],
                    h[r * 8 + 1], h[r * 8 + 2], h[r * 8 + 3], h[r * 8 + 4]);
        return 0;
    }
    // training forwards (stash != NULL) take the split launch only when the option is 2: it costs two small launches
    // per step, which matter in a data-parallel step of 512 rays per GPU, and the loss does not need it
    if (stash && ctx->precise_last < 2) return 0;
    return tc_precise_last(ctx, which, half, B, S, ro, rd, t, sigma, stash, st);
}

}  // namespace nb
