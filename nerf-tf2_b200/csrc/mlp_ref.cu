// NERFB200_FP32: the exact-arithmetic check path of the 8x256 MLP (core/model.py:334-394) on CUDA
// cores -- layer-by-layer fp32 FMA GEMMs with the activations in HBM, mirroring the op-by-op
// structure of the reference graph (MatMul + BiasAdd + Relu/Sigmoid).
// It exists so that (i) every other kernel can be parity-checked end to end on the GPU with
// fp32 network outputs (searchsorted indices, rendered pixels) and (ii) the tensor-core kernel
// (mlp_tc.cu) has an on-device fp32 reference at full problem sizes. It is NOT the fast path.
#include "common.cuh"
#include "mlp.cuh"

namespace nb {

// ------------------------------------------------------------------------------------------
// Generic strided SGEMM: C[m,n] (+)= sum_k A(m,k)*B(k,n), A(m,k)=A[m*sAm+k*sAk], B(k,n)=B[k*sBk+n*sBn].
// 64x64 tile, BK=16, 256 threads, 4x4 outputs per thread. grid.z splits K (atomicAdd epilogue).
// Epilogue: optional bias[n] and activation (0 none, 1 relu, 2 sigmoid); accumulate adds into C.
constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int64_t sAm, int64_t sAk, const float* __restrict__ B,
             int64_t sBk, int64_t sBn, float* __restrict__ C, int64_t ldc, const float* __restrict__ bias, int act,
             int accumulate, int k_per_split) {
    __shared__ float sA[BK][BM + 4];
    __shared__ float sB[BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);
    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4] = {};
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
        for (int i = 0; i < (BM * BK) / 256; ++i) {
            int idx = tid + i * 256;
            int mm, kk;
            if (sAk == 1) { kk = idx % BK; mm = idx / BK; } else { mm = idx % BM; kk = idx / BM; }
            int m = m0 + mm, k = k0 + kk;
            sA[kk][mm] = (m < M && k < kend) ? A[(int64_t)m * sAm + (int64_t)k * sAk] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < (BN * BK) / 256; ++i) {
            int idx = tid + i * 256;
            int nn, kk;
            if (sBn == 1) { nn = idx % BN; kk = idx / BN; } else { kk = idx % BK; nn = idx / BK; }
            int n = n0 + nn, k = k0 + kk;
            sB[kk][nn] = (n < N && k < kend) ? B[(int64_t)k * sBk + (int64_t)n * sBn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            float* c = C + (int64_t)m * ldc + n;
            if (gridDim.z > 1) { atomicAdd(c, v); continue; }
            if (accumulate) v += *c;
            if (bias) v += bias[n];
            if (act == 1) v = fmaxf(v, 0.f);
            else if (act == 2) v = 1.f / (1.f + expf(-v));
            *c = v;
        }
    }
}

static int gemm(cudaStream_t st, int M, int N, int K, const float* A, int64_t sAm, int64_t sAk, const float* B,
                int64_t sBk, int64_t sBn, float* C, int64_t ldc, const float* bias, int act, int accumulate,
                int splits = 1) {
    if (M == 0) return 0;
    int kps = (K + splits - 1) / splits;
    kps = (kps + BK - 1) / BK * BK;
    splits = (K + kps - 1) / kps;
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, splits);
    sgemm_kernel<<<grid, 256, 0, st>>>(M, N, K, A, sAm, sAk, B, sBk, sBn, C, ldc, bias, act, accumulate, kps);
    NB_LAUNCH_CHECK();
    return 0;
}

// bias + activation applied in place (second pass of a two-source layer)
__global__ void bias_act_kernel(int64_t n_elem, int N, float* __restrict__ Y, const float* __restrict__ bias, int act) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_elem) return;
    float v = Y[i] + bias[i % N];
    if (act == 1) v = fmaxf(v, 0.f);
    else if (act == 2) v = 1.f / (1.f + expf(-v));
    Y[i] = v;
}

// encoders straight from rays: xyz = o + t*d (separate mul/add, utils/ray_utils.py:251), dir = d
__global__ void encode_rows_kernel(int64_t row0, int64_t R, int S, const float* __restrict__ ro,
                                   const float* __restrict__ rd, const float* __restrict__ t,
                                   float* __restrict__ enc_xyz, float* __restrict__ enc_dir) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * 90) return;
    int64_t r = i / 90;
    int j = (int)(i - r * 90);
    int64_t row = row0 + r, ray = row / S;
    bool is_dir = j >= 63;
    int jj = is_dir ? j - 63 : j;
    int L = is_dir ? 4 : 10;
    int d, l = 0, s = 0;
    if (jj < 3) d = jj;
    else { int q = jj - 3; d = q / (2 * L); l = (q - d * 2 * L) >> 1; s = q & 1; }
    float dv = rd[3 * ray + d];
    float x = is_dir ? dv : __fadd_rn(ro[3 * ray + d], __fmul_rn(t[row], dv));
    float v = x;
    if (jj >= 3) {
        float e = __fmul_rn(x, __fmul_rn((float)(1 << l), 3.14159274101257324f));
        v = s ? cosf(e) : sinf(e);
    }
    if (is_dir) enc_dir[r * 27 + jj] = v; else enc_xyz[r * 63 + jj] = v;
}

// dZ = dY * act'(Y) in place on dY (act: 1 relu, 2 sigmoid)
__global__ void act_grad_kernel(int64_t n, float* __restrict__ dY, const float* __restrict__ Y, int act) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float y = Y[i];
    dY[i] = act == 1 ? (y > 0.f ? dY[i] : 0.f) : dY[i] * y * (1.f - y);
}

// db[n] += sum_r dZ[r,n]
__global__ void colsum_kernel(int64_t R, int N, const float* __restrict__ dZ, float* __restrict__ db) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    int64_t r0 = (int64_t)blockIdx.y * 1024, r1 = min(R, r0 + 1024);
    float s = 0.f;
    for (int64_t r = r0; r < r1; ++r) s += dZ[r * N + n];
    atomicAdd(db + n, s);
}

__global__ void add_inplace_kernel(int64_t n, float* __restrict__ a, const float* __restrict__ b) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += b[i];
}

static inline unsigned nblk(int64_t n) { return (unsigned)((n + 255) / 256); }

// Stash layout for R rows (fp32): enc_xyz[R,63] enc_dir[R,27] h[8][R,256] bott[R,256] h9[R,128] rgb[R,3] sigma[R]
struct RefStash {
    float *enc_xyz, *enc_dir, *h[8], *bott, *h9, *rgb, *sigma;
    static int64_t floats(int64_t R) { return R * (63 + 27 + 8 * 256 + 256 + 128 + 3 + 1); }
    void bind(float* p, int64_t R) {
        enc_xyz = p; p += R * 63;
        enc_dir = p; p += R * 27;
        for (int i = 0; i < 8; ++i) { h[i] = p; p += R * 256; }
        bott = p; p += R * 256;
        h9 = p; p += R * 128;
        rgb = p; p += R * 3;
        sigma = p;
    }
};

int64_t ref_workspace_bytes(int64_t R, int training) {
    if (training) return (int64_t)R * (256 + 256 + 128 + 3 + 1) * 4;  // backward scratch
    int64_t ch = R < kRefChunkRows ? R : kRefChunkRows;
    return RefStash::floats(ch) * 4;
}
int64_t ref_stash_bytes(int64_t R) { return RefStash::floats(R) * 4; }

static int ref_forward_rows(cudaStream_t st, const float* P, int64_t row0, int64_t R, int S, const float* ro,
                            const float* rd, const float* t, RefStash& a, float* rgb_out, float* sigma_out) {
    encode_rows_kernel<<<nblk(R * 90), 256, 0, st>>>(row0, R, S, ro, rd, t, a.enc_xyz, a.enc_dir);
    NB_LAUNCH_CHECK();
    int rc;
    const int M = (int)R;
    // dense_0..4
    if ((rc = gemm(st, M, 256, 63, a.enc_xyz, 63, 1, P + kernel_offset(L0), 256, 1, a.h[0], 256, P + bias_offset(L0), 1, 0))) return rc;
    for (int i = 1; i <= 4; ++i)
        if ((rc = gemm(st, M, 256, 256, a.h[i - 1], 256, 1, P + kernel_offset(i), 256, 1, a.h[i], 256, P + bias_offset(i), 1, 0))) return rc;
    // dense_5 on concat[h4, enc_xyz]
    if ((rc = gemm(st, M, 256, 256, a.h[4], 256, 1, P + kernel_offset(L5), 256, 1, a.h[5], 256, nullptr, 0, 0))) return rc;
    if ((rc = gemm(st, M, 256, 63, a.enc_xyz, 63, 1, P + kernel_offset(L5) + 256 * 256, 256, 1, a.h[5], 256, P + bias_offset(L5), 1, 1))) return rc;
    for (int i = 6; i <= 7; ++i)
        if ((rc = gemm(st, M, 256, 256, a.h[i - 1], 256, 1, P + kernel_offset(i), 256, 1, a.h[i], 256, P + bias_offset(i), 1, 0))) return rc;
    if ((rc = gemm(st, M, 1, 256, a.h[7], 256, 1, P + kernel_offset(LSIGMA), 1, 1, sigma_out, 1, P + bias_offset(LSIGMA), 1, 0))) return rc;
    if ((rc = gemm(st, M, 256, 256, a.h[7], 256, 1, P + kernel_offset(L8), 256, 1, a.bott, 256, P + bias_offset(L8), 0, 0))) return rc;
    // dense_9 on concat[bott, enc_dir]
    if ((rc = gemm(st, M, 128, 256, a.bott, 256, 1, P + kernel_offset(L9), 128, 1, a.h9, 128, nullptr, 0, 0))) return rc;
    if ((rc = gemm(st, M, 128, 27, a.enc_dir, 27, 1, P + kernel_offset(L9) + 256 * 128, 128, 1, a.h9, 128, P + bias_offset(L9), 1, 1))) return rc;
    if ((rc = gemm(st, M, 3, 128, a.h9, 128, 1, P + kernel_offset(LRGB), 3, 1, rgb_out, 3, P + bias_offset(LRGB), 2, 0))) return rc;
    return 0;
}

int ref_forward(cudaStream_t st, const float* P, int64_t B, int S, const float* ro, const float* rd, const float* t,
                float* rgb, float* sigma, void* workspace, void* stash) {
    const int64_t R = B * S;
    RefStash a;
    if (stash) {
        // training: one pass over all rows, activations kept
        NB_CHECK_ARG(R < (int64_t)1 << 31, "mlp_forward(fp32, training): too many rows");
        a.bind((float*)stash, R);
        int rc = ref_forward_rows(st, P, 0, R, S, ro, rd, t, a, a.rgb, a.sigma);
        if (rc) return rc;
        NB_CUDA(cudaMemcpyAsync(rgb, a.rgb, R * 3 * 4, cudaMemcpyDeviceToDevice, st));
        NB_CUDA(cudaMemcpyAsync(sigma, a.sigma, R * 4, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    NB_CHECK_ARG(workspace != nullptr, "mlp_forward(fp32): workspace required");
    for (int64_t r0 = 0; r0 < R; r0 += kRefChunkRows) {
        int64_t n = R - r0 < kRefChunkRows ? R - r0 : kRefChunkRows;
        a.bind((float*)workspace, R < kRefChunkRows ? R : kRefChunkRows);
        int rc = ref_forward_rows(st, P, r0, n, S, ro, rd, t, a, rgb + 3 * r0, sigma + r0);
        if (rc) return rc;
    }
    return 0;
}

// Backward through the stashed activations. workspace: dA[R,256], dB[R,256], d9[R,128], dzr[R,3], dzs[R].
int ref_backward(cudaStream_t st, const float* P, int64_t B, int S, const float* d_rgb, const float* d_sigma,
                 float* G, void* workspace, void* stash) {
    const int64_t R64 = B * S;
    NB_CHECK_ARG(workspace && stash, "mlp_backward(fp32): workspace and stash required");
    NB_CHECK_ARG(R64 < (int64_t)1 << 31, "mlp_backward(fp32): too many rows");
    const int R = (int)R64;
    RefStash a;
    a.bind((float*)stash, R);
    float* dA = (float*)workspace;
    float* dB = dA + (int64_t)R * 256;
    float* d9 = dB + (int64_t)R * 256;
    float* dzr = d9 + (int64_t)R * 128;
    float* dzs = dzr + (int64_t)R * 3;
    int rc;
    const int splits = R > 4096 ? 64 : 8;   // split the row reduction of dW across blocks
    auto dW = [&](int M, int N, const float* A, int lda, const float* dZ, float* Gk) {
        // Gk[M,N] += A[R,M]^T . dZ[R,N]
        return gemm(st, M, N, R, A, 1, lda, dZ, N, 1, Gk, N, nullptr, 0, 1, splits);
    };
    auto dbias = [&](int N, const float* dZ, float* Gb) {
        dim3 grid((N + 127) / 128, (R + 1023) / 1024);
        colsum_kernel<<<grid, 128, 0, st>>>(R, N, dZ, Gb);
        return (int)cudaGetLastError();
    };
    // rgb head: dZ = d_rgb * rgb*(1-rgb)
    NB_CUDA(cudaMemcpyAsync(dzr, d_rgb, (int64_t)R * 3 * 4, cudaMemcpyDeviceToDevice, st));
    act_grad_kernel<<<nblk((int64_t)R * 3), 256, 0, st>>>((int64_t)R * 3, dzr, a.rgb, 2);
    if ((rc = dW(128, 3, a.h9, 128, dzr, G + kernel_offset(LRGB)))) return rc;
    if ((rc = dbias(3, dzr, G + bias_offset(LRGB)))) return rc;
    // d h9 = dzr . Wrgb^T ; relu
    if ((rc = gemm(st, R, 128, 3, dzr, 3, 1, P + kernel_offset(LRGB), 1, 3, d9, 128, nullptr, 0, 0))) return rc;
    act_grad_kernel<<<nblk((int64_t)R * 128), 256, 0, st>>>((int64_t)R * 128, d9, a.h9, 1);
    if ((rc = dW(256, 128, a.bott, 256, d9, G + kernel_offset(L9)))) return rc;
    if ((rc = dW(27, 128, a.enc_dir, 27, d9, G + kernel_offset(L9) + 256 * 128))) return rc;
    if ((rc = dbias(128, d9, G + bias_offset(L9)))) return rc;
    // d bott = d9 . W9[0:256]^T  (dense_8 is linear)
    if ((rc = gemm(st, R, 256, 128, d9, 128, 1, P + kernel_offset(L9), 1, 128, dA, 256, nullptr, 0, 0))) return rc;
    if ((rc = dW(256, 256, a.h[7], 256, dA, G + kernel_offset(L8)))) return rc;
    if ((rc = dbias(256, dA, G + bias_offset(L8)))) return rc;
    // d h7 = dbott . W8^T + dzs . Wsigma^T
    if ((rc = gemm(st, R, 256, 256, dA, 256, 1, P + kernel_offset(L8), 1, 256, dB, 256, nullptr, 0, 0))) return rc;
    NB_CUDA(cudaMemcpyAsync(dzs, d_sigma, (int64_t)R * 4, cudaMemcpyDeviceToDevice, st));
    act_grad_kernel<<<nblk(R), 256, 0, st>>>(R, dzs, a.sigma, 1);
    if ((rc = dW(256, 1, a.h[7], 256, dzs, G + kernel_offset(LSIGMA)))) return rc;
    if ((rc = dbias(1, dzs, G + bias_offset(LSIGMA)))) return rc;
    if ((rc = gemm(st, R, 256, 1, dzs, 1, 1, P + kernel_offset(LSIGMA), 1, 1, dB, 256, nullptr, 0, 1))) return rc;
    // dense_7 .. dense_0; cur = gradient w.r.t. h[i] (post-activation)
    float* cur = dB;
    float* nxt = dA;
    for (int i = 7; i >= 0; --i) {
        act_grad_kernel<<<nblk((int64_t)R * 256), 256, 0, st>>>((int64_t)R * 256, cur, a.h[i], 1);
        if ((rc = dbias(256, cur, G + bias_offset(i)))) return rc;
        if (i == 0) {
            if ((rc = dW(63, 256, a.enc_xyz, 63, cur, G + kernel_offset(L0)))) return rc;
            break;
        }
        if ((rc = dW(256, 256, a.h[i - 1], 256, cur, G + kernel_offset(i)))) return rc;
        if (i == 5)
            if ((rc = dW(63, 256, a.enc_xyz, 63, cur, G + kernel_offset(L5) + 256 * 256))) return rc;
        if ((rc = gemm(st, R, 256, 256, cur, 256, 1, P + kernel_offset(i), 1, 256, nxt, 256, nullptr, 0, 0))) return rc;
        float* tmp = cur; cur = nxt; nxt = tmp;
    }
    NB_LAUNCH_CHECK();
    return 0;
}

}  // namespace nb
