// HBM layouts shared by the training kernels of the tensor-core path.
//
// Everything is organised per 128-row tile and, inside a tile, as [128 x 64] "chunk images":
// 16 KB blocks with exactly the byte layout the UMMA operands have in shared memory (128 B per row,
// 16-byte units XOR-swizzled by row & 7). A chunk image moves between HBM and shared memory with one
// bulk-TMA copy and is directly usable as a K-major operand (K = features: forward / backward-data)
// or as an MN-major operand (K = rows: weight gradients) without any transposition.
#pragma once
#include <stdint.h>

namespace nb {

// ---- activation stash written by the training forward (per tile)
constexpr int kStashChunkEncXyz = 0;                                  // enc_xyz (63 + 1 zero column)
__host__ __device__ constexpr int stash_chunk_Y(int l) { return 1 + 4 * l; }   // post-ReLU output of dense_l, l = 0..7
constexpr int kStashChunkBott = 33;                                   // dense_8 output (linear), 4 chunks
constexpr int kStashChunkEncDir = 37;                                 // enc_dir (27 + zero columns)
constexpr int kStashChunkY9 = 38;                                     // post-ReLU output of dense_9 (128 wide), 2 chunks
constexpr int kStashChunks = 40;
constexpr int kStashMaskLayers = 9;                                   // ReLU bitmasks: Y0..Y7 -> 0..7, Y9 -> 8
constexpr int kStashMaskOfs = kStashChunks * 16384;                   // [layer][row][8 x u32]
// ReLU bitmask word of a 32-column group: column i of the group lives at bit mask_bit(i) -- even columns in
// bits 0..15, odd columns in bits 16..31 -- so that the forward epilogue builds it from the packed 16-bit
// pairs with one compare + one LOP3 per pair (word k = columns 2k, 2k+1 -> bits k and 16+k).
__host__ __device__ constexpr int mask_bit(int i) { return ((i & 1) << 4) | (i >> 1); }
constexpr int kStashOutOfs = kStashMaskOfs + kStashMaskLayers * 128 * 32;   // rgb[128][3] fp32, then sigma[128] fp32
constexpr int kStashTileBytes = kStashOutOfs + 128 * 16;
static_assert(kStashTileBytes % 16 == 0, "tile stash must keep 16-byte alignment");

// ---- gradient stash written by the backward-data kernel (per tile): dZ = dLoss/d(pre-activation)
constexpr int kGradChunkZ9 = 0;                                       // 2 chunks (128 wide)
constexpr int kGradChunkZ8 = 2;                                       // 4 chunks
__host__ __device__ constexpr int grad_chunk_Z(int l) { return 6 + 4 * (7 - l); }   // dense_l, l = 7..0
constexpr int kGradChunkHead = 38;                                    // cols 0..2 = dZ_rgb, col 3 = dZ_sigma, rest 0
constexpr int kGradChunks = 39;
constexpr int kGradTileBytes = kGradChunks * 16384;

}  // namespace nb
