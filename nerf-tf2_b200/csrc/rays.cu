// Camera rays, stratified sampler, network-input materialisation, positional encoding,
// depth-map type_2. All HBM-bound elementwise kernels: one thread per output element group,
// coalesced stores, grid sized by the element count.
#include "common.cuh"

#include <stdarg.h>
#include <atomic>

namespace nb {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

struct CamD { double K[9]; double c2w[16]; };
struct CamF { float K[9]; float c2w[16]; };

// a1: get_rays (utils/ray_utils.py:6-51): fp64 maths, result cast to fp32.
__global__ void get_rays_f64_kernel(int W, CamD cam, int64_t ray0, int64_t n, float* __restrict__ ro,
                                    float* __restrict__ rd) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t id = ray0 + i;
    double u = (double)(id % W), v = (double)(id / W);
    double x = (u - cam.K[2]) / cam.K[0];
    double y = (v - cam.K[5]) / cam.K[4];
    double z = 1.0;
    double d0 = cam.c2w[0] * x + cam.c2w[1] * y + cam.c2w[2] * z;
    double d1 = cam.c2w[4] * x + cam.c2w[5] * y + cam.c2w[6] * z;
    double d2 = cam.c2w[8] * x + cam.c2w[9] * y + cam.c2w[10] * z;
    double mag = sqrt(d0 * d0 + d1 * d1 + d2 * d2) + 1e-8;
    rd[3 * i + 0] = (float)(d0 / mag);
    rd[3 * i + 1] = (float)(d1 / mag);
    rd[3 * i + 2] = (float)(d2 / mag);
    ro[3 * i + 0] = (float)cam.c2w[3];
    ro[3 * i + 1] = (float)cam.c2w[7];
    ro[3 * i + 2] = (float)cam.c2w[11];
}

// a2: get_rays_tf (utils/ray_utils.py:53-106): fp32 maths.
__device__ __forceinline__ void ray_f32(const CamF& cam, int W, int64_t id, float* ro, float* rd) {
    float u = (float)(id % W), v = (float)(id / W);
    float x = __fdiv_rn(__fsub_rn(u, cam.K[2]), cam.K[0]);
    float y = __fdiv_rn(__fsub_rn(v, cam.K[5]), cam.K[4]);
    float z = 1.0f;
    // directions @ R^T: sum over k of dir[k] * R[j][k], accumulated left to right.
    float d[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float a = __fmul_rn(x, cam.c2w[4 * j + 0]);
        a = __fadd_rn(a, __fmul_rn(y, cam.c2w[4 * j + 1]));
        a = __fadd_rn(a, __fmul_rn(z, cam.c2w[4 * j + 2]));
        d[j] = a;
    }
    float s = __fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2]));
    float mag = __fadd_rn(__fsqrt_rn(s), 1e-8f);
    rd[0] = __fdiv_rn(d[0], mag);
    rd[1] = __fdiv_rn(d[1], mag);
    rd[2] = __fdiv_rn(d[2], mag);
    ro[0] = cam.c2w[3];
    ro[1] = cam.c2w[7];
    ro[2] = cam.c2w[11];
}

__global__ void get_rays_f32_kernel(int W, CamF cam, int64_t ray0, int64_t n, const int32_t* __restrict__ ids,
                                    float* __restrict__ ro, float* __restrict__ rd) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t id = ids ? (int64_t)ids[i] : ray0 + i;
    float o[3], d[3];
    ray_f32(cam, W, id, o, d);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        ro[3 * i + j] = o[j];
        rd[3 * i + j] = d[j];
    }
}

// a3: bin edge k of tf.linspace semantics: endpoints exact, interior
// start + delta*k with one multiply and one add.
__device__ __forceinline__ float bin_edge(float near, float far, int k, int Nc, int lin_inv) {
    if (!lin_inv) {
        if (k == 0) return near;
        if (k == Nc) return far;
        float delta = __fdiv_rn(__fsub_rn(far, near), (float)Nc);
        return __fadd_rn(near, __fmul_rn(delta, (float)k));
    }
    float a = __fdiv_rn(1.0f, near), b = __fdiv_rn(1.0f, far);
    float v;
    if (k == 0) v = a;
    else if (k == Nc) v = b;
    else {
        float delta = __fdiv_rn(__fsub_rn(b, a), (float)Nc);
        v = __fadd_rn(a, __fmul_rn(delta, (float)k));
    }
    return __fdiv_rn(1.0f, v);
}

// One thread per (ray, group of 4 consecutive bins): the ray's reciprocals and the linspace delta are computed once
// per group instead of once per edge, the 5 edges of the group are shared by its 4 samples, and ONE Philox call
// yields the 4 uniforms (counter block = k >> 2, as everywhere). Every value is produced by exactly the operations of
// bin_edge() above, in the same order: bit-identical to the one-thread-per-edge form.
__global__ void sample_coarse_kernel(int64_t B, int Nc, int lin_inv, int perturb, const float* __restrict__ near,
                                     const float* __restrict__ far, const float* __restrict__ u_in, uint64_t seed,
                                     const int64_t* __restrict__ step_dev, int64_t ray0, float* __restrict__ t_vals,
                                     float* __restrict__ edges) {
    if (step_dev) seed ^= (uint64_t)__ldg(step_dev + 1);      // device-resident step counter (CUDA-graph replays)
    const int G = (Nc + 3) >> 2;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * G) return;
    const int64_t ray = i / G;
    const int k0 = 4 * (int)(i - ray * G);
    const float n = near[ray], f = far[ray];
    const float a = lin_inv ? __fdiv_rn(1.0f, n) : n, b = lin_inv ? __fdiv_rn(1.0f, f) : f;
    const float delta = __fdiv_rn(__fsub_rn(b, a), (float)Nc);
    float e[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int k = k0 + j;
        float v = (k == 0) ? a : ((k >= Nc) ? b : __fadd_rn(a, __fmul_rn(delta, (float)k)));
        e[j] = lin_inv ? __fdiv_rn(1.0f, v) : v;
    }
    float* erow = edges + ray * (Nc + 1);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (k0 + j <= Nc) erow[k0 + j] = e[j];
    if (k0 + 4 == Nc) erow[Nc] = e[4];
    float u[4] = {0.f, 0.f, 0.f, 0.f};
    if (perturb) {
        if (u_in) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (k0 + j < Nc) u[j] = u_in[ray * Nc + k0 + j];
        } else {
            const float4 r = philox_uniform4(seed, (uint64_t)(ray0 + ray), (uint32_t)(k0 >> 2), 0u);
            u[0] = r.x; u[1] = r.y; u[2] = r.z; u[3] = r.w;
        }
    }
    float t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        t[j] = perturb ? __fadd_rn(e[j], __fmul_rn(u[j], __fsub_rn(e[j + 1], e[j])))     // utils/ray_utils.py:234
                       : __fmul_rn(0.5f, __fadd_rn(e[j], e[j + 1]));                      // utils/ray_utils.py:243
    float* trow = t_vals + ray * Nc + k0;
    if ((Nc & 3) == 0 && (((uintptr_t)t_vals) & 15) == 0) {
        *reinterpret_cast<float4*>(trow) = make_float4(t[0], t[1], t[2], t[3]);
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (k0 + j < Nc) trow[j] = t[j];
    }
}

// xyz = o + t*d, dirs broadcast (utils/ray_utils.py:251-258). One thread per row.
__global__ void make_inputs_kernel(int64_t R, int S, const float* __restrict__ ro, const float* __restrict__ rd,
                                   const float* __restrict__ t, float* __restrict__ xyz, float* __restrict__ dirs) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int64_t ray = r / S;
    float tt = t[r];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float d = rd[3 * ray + j];
        xyz[3 * r + j] = __fadd_rn(ro[3 * ray + j], __fmul_rn(tt, d));
        dirs[3 * r + j] = d;
    }
}

// a4: PositionalEncoder.call (core/model.py:305-332). One thread per (row, output feature).
__global__ void posenc_kernel(int64_t R, int L, const float* __restrict__ x, float* __restrict__ out) {
    int F = 3 + 6 * L;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * F) return;
    int64_t r = i / F;
    int j = (int)(i - r * F);
    float v;
    if (j < 3) v = x[3 * r + j];
    else {
        int q = j - 3;
        int d = q / (2 * L);
        int l = (q - d * 2 * L) >> 1;
        int s = q & 1;
        float m = __fmul_rn((float)(1 << l), 3.14159274101257324f);   // fl32(2^l) * fl32(pi)
        float e = __fmul_rn(x[3 * r + d], m);
        v = s ? cosf(e) : sinf(e);
    }
    out[i] = v;
}

// a15: create_depth_map type_2 (utils/ray_utils.py:122-130), fp64 maths.
__global__ void depth_type2_kernel(int W, int64_t n, CamD cam, CamD inv, double inv_scale,
                                   const float* __restrict__ depth, float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double u = (double)(i % W), v = (double)(i / W);
    double x = (u - cam.K[2]) / cam.K[0], y = (v - cam.K[5]) / cam.K[4], z = 1.0;
    double d0 = cam.c2w[0] * x + cam.c2w[1] * y + cam.c2w[2] * z;
    double d1 = cam.c2w[4] * x + cam.c2w[5] * y + cam.c2w[6] * z;
    double d2 = cam.c2w[8] * x + cam.c2w[9] * y + cam.c2w[10] * z;
    double mag = sqrt(d0 * d0 + d1 * d1 + d2 * d2) + 1e-8;
    double dep = (double)depth[i] * inv_scale;
    double p0 = cam.c2w[3] + d0 / mag * dep, p1 = cam.c2w[7] + d1 / mag * dep, p2 = cam.c2w[11] + d2 / mag * dep;
    out[i] = (float)(inv.c2w[8] * p0 + inv.c2w[9] * p1 + inv.c2w[10] * p2 + inv.c2w[11]);
}


// ---- callers either side of the path ---------------------------------------
// next-2: sample-mode training input (core/base_dataset.py:555-621): B uniform pixel ids in [0, n_pixels)
// (Philox, like tf.random.uniform(..., dtype=int32)), and the gather of their colours from a uint8 image.
__global__ void sample_pixels_kernel(int64_t B, uint32_t n_pixels, uint64_t seed, uint64_t step, int32_t* __restrict__ ids) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    uint4 r = philox4x32_10(make_uint4((uint32_t)(i >> 2), (uint32_t)((uint64_t)(i >> 2) >> 32), (uint32_t)step, 2u),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    uint32_t x = (i & 3) == 0 ? r.x : (i & 3) == 1 ? r.y : (i & 3) == 2 ? r.z : r.w;
    ids[i] = (int32_t)(((uint64_t)x * (uint64_t)n_pixels) >> 32);        // unbiased enough for n_pixels << 2^32
}
__global__ void gather_rgb_u8_kernel(int64_t B, const uint8_t* __restrict__ img, const int32_t* __restrict__ ids,
                                     float* __restrict__ rgb) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 3) return;
    int64_t r = i / 3;
    int c = (int)(i - r * 3);
    rgb[i] = __fdiv_rn((float)img[(int64_t)ids[r] * 3 + c], 255.0f);      // cast then / 255.0 (:596-598)
}
// next-3: per-view post-processing of main/eval.py:53-64 and main/render.py:96-100 on the device:
// img_u8 = uint8(clip(pred*255, 0, 255)); sq_err += sum((gt/255 - clip(pred*255,0,255)/255)^2) for PSNR.
__global__ void postprocess_rgb_kernel(int64_t n, const float* __restrict__ pred, const uint8_t* __restrict__ gt,
                                       uint8_t* __restrict__ out, double* __restrict__ sq_err) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float e = 0.f;
    if (i < n) {
        float v = fminf(fmaxf(__fmul_rn(pred[i], 255.0f), 0.0f), 255.0f);
        if (out) out[i] = (uint8_t)v;                                     // astype(np.uint8) truncates
        if (gt) {
            float d = __fsub_rn(__fdiv_rn((float)gt[i], 255.0f), __fdiv_rn(v, 255.0f));
            e = d * d;
        }
    }
    if (sq_err && gt) {
        e = warp_sum(e);
        if ((threadIdx.x & 31) == 0 && e != 0.f) atomicAdd(sq_err, (double)e);
    }
}

static inline unsigned blocks_for(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace nb

using namespace nb;

extern "C" {

const char* nerfb200_last_error(void) { return nb::g_err; }
int64_t nerfb200_launch_count(void) { return (int64_t)nb::g_launches.load(); }
int nerfb200_abi_version(void) { return NERFB200_ABI_VERSION; }

int nerfb200_param_offsets(int64_t* offsets) {
    NB_CHECK_ARG(offsets != nullptr, "param_offsets: NULL output");
    for (int l = 0; l < kNumLayers; ++l) {
        offsets[2 * l] = kernel_offset(l);
        offsets[2 * l + 1] = bias_offset(l);
    }
    offsets[2 * kNumLayers] = kParamsPerModel;
    return 0;
}

int nerfb200_get_rays(int H, int W, const double* K, const double* c2w, int64_t ray0, int64_t n_rays,
                      float* rays_o, float* rays_d, void* stream) {
    NB_CHECK_ARG(H > 0 && W > 0 && K && c2w && rays_o && rays_d, "get_rays: bad arguments");
    NB_CHECK_ARG(ray0 >= 0 && n_rays >= 0 && ray0 + n_rays <= (int64_t)H * W, "get_rays: ray range outside image");
    if (n_rays == 0) return 0;
    CamD cam;
    for (int i = 0; i < 9; ++i) cam.K[i] = K[i];
    for (int i = 0; i < 16; ++i) cam.c2w[i] = c2w[i];
    get_rays_f64_kernel<<<blocks_for(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(W, cam, ray0, n_rays, rays_o, rays_d);
    NB_LAUNCH_CHECK();
    return 0;
}

static int get_rays_f32_impl(int H, int W, const float* K, const float* c2w, int64_t ray0, int64_t n_rays,
                             const int32_t* ids, float* rays_o, float* rays_d, void* stream) {
    NB_CHECK_ARG(H > 0 && W > 0 && K && c2w && rays_o && rays_d, "get_rays_f32: bad arguments");
    NB_CHECK_ARG(n_rays >= 0, "get_rays_f32: negative ray count");
    if (n_rays == 0) return 0;
    CamF cam;
    for (int i = 0; i < 9; ++i) cam.K[i] = K[i];
    for (int i = 0; i < 16; ++i) cam.c2w[i] = c2w[i];
    get_rays_f32_kernel<<<blocks_for(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(W, cam, ray0, n_rays, ids, rays_o, rays_d);
    NB_LAUNCH_CHECK();
    return 0;
}

int nerfb200_get_rays_f32(int H, int W, const float* K, const float* c2w, int64_t ray0, int64_t n_rays,
                          float* rays_o, float* rays_d, void* stream) {
    NB_CHECK_ARG(ray0 >= 0 && ray0 + n_rays <= (int64_t)H * W, "get_rays_f32: ray range outside image");
    return get_rays_f32_impl(H, W, K, c2w, ray0, n_rays, nullptr, rays_o, rays_d, stream);
}

int nerfb200_get_rays_at(int H, int W, const float* K, const float* c2w, const int32_t* pixel_ids, int64_t n_rays,
                         float* rays_o, float* rays_d, void* stream) {
    NB_CHECK_ARG(pixel_ids != nullptr, "get_rays_at: NULL pixel ids");
    return get_rays_f32_impl(H, W, K, c2w, 0, n_rays, pixel_ids, rays_o, rays_d, stream);
}

int nerfb200_sample_coarse(int64_t B, int Nc, int lin_inv_depth, int perturb, const float* near, const float* far,
                           const float* u_coarse, uint64_t seed, const int64_t* step_state, int64_t ray0, float* t_vals,
                           float* bin_edges, void* stream) {
    NB_CHECK_ARG(B >= 0 && Nc >= 2, "sample_coarse: bad shape");
    if (B == 0) return 0;
    NB_CHECK_ARG(near && far && t_vals && bin_edges, "sample_coarse: NULL pointer");
    sample_coarse_kernel<<<blocks_for(B * ((Nc + 3) / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        B, Nc, lin_inv_depth, perturb, near, far, u_coarse, seed, step_state, ray0, t_vals, bin_edges);
    NB_LAUNCH_CHECK();
    return 0;
}

int nerfb200_make_inputs(int64_t B, int S, const float* rays_o, const float* rays_d, const float* t_vals, float* xyz,
                         float* dirs, void* stream) {
    NB_CHECK_ARG(B >= 0 && S > 0, "make_inputs: bad shape");
    if (B == 0) return 0;
    NB_CHECK_ARG(rays_o && rays_d && t_vals && xyz && dirs, "make_inputs: NULL pointer");
    make_inputs_kernel<<<blocks_for(B * S, 256), 256, 0, (cudaStream_t)stream>>>(B * S, S, rays_o, rays_d, t_vals, xyz, dirs);
    NB_LAUNCH_CHECK();
    return 0;
}

int nerfb200_positional_encode(int64_t R, int L, const float* x, float* out, void* stream) {
    NB_CHECK_ARG(R >= 0 && L >= 1 && L <= 16, "positional_encode: bad shape");
    if (R == 0) return 0;
    NB_CHECK_ARG(x && out, "positional_encode: NULL pointer");
    posenc_kernel<<<blocks_for(R * (3 + 6 * L), 256), 256, 0, (cudaStream_t)stream>>>(R, L, x, out);
    NB_LAUNCH_CHECK();
    return 0;
}


int nerfb200_sample_pixels(int64_t B, int64_t n_pixels, uint64_t seed, uint64_t step, int32_t* pixel_ids, void* stream) {
    NB_CHECK_ARG(B >= 0 && n_pixels > 0 && n_pixels < ((int64_t)1 << 31), "sample_pixels: bad shape");
    if (B == 0) return 0;
    NB_CHECK_ARG(pixel_ids != nullptr, "sample_pixels: NULL output");
    sample_pixels_kernel<<<blocks_for(B, 256), 256, 0, (cudaStream_t)stream>>>(B, (uint32_t)n_pixels, seed, step, pixel_ids);
    NB_LAUNCH_CHECK();
    return 0;
}

int nerfb200_gather_rgb_u8(int64_t B, const uint8_t* image, const int32_t* pixel_ids, float* rgb, void* stream) {
    NB_CHECK_ARG(B >= 0, "gather_rgb_u8: bad shape");
    if (B == 0) return 0;
    NB_CHECK_ARG(image && pixel_ids && rgb, "gather_rgb_u8: NULL pointer");
    gather_rgb_u8_kernel<<<blocks_for(B * 3, 256), 256, 0, (cudaStream_t)stream>>>(B, image, pixel_ids, rgb);
    NB_LAUNCH_CHECK();
    return 0;
}

int nerfb200_postprocess_rgb(int64_t n_values, const float* pred_rgb, const uint8_t* gt_u8, uint8_t* out_u8, double* sq_err,
                             void* stream) {
    NB_CHECK_ARG(n_values >= 0, "postprocess_rgb: bad shape");
    if (n_values == 0) return 0;
    NB_CHECK_ARG(pred_rgb != nullptr && (out_u8 || (gt_u8 && sq_err)), "postprocess_rgb: nothing to do");
    postprocess_rgb_kernel<<<blocks_for(n_values, 256), 256, 0, (cudaStream_t)stream>>>(n_values, pred_rgb, gt_u8, out_u8, sq_err);
    NB_LAUNCH_CHECK();
    return 0;
}

static void invert_rigid4(const double* m, double* inv) {
    // general 4x4 inverse via cofactors is overkill: poses are [sR | t; 0 0 0 1] with uniform scale
    // folded into t only (reconfigure_scene_scale scales the whole top 3x4 block), so use a plain
    // Gauss-Jordan on the 4x4 to stay faithful to np.linalg.inv.
    double a[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) { a[i][j] = m[4 * i + j]; a[i][4 + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < 4; ++c) {
        int p = c;
        for (int r = c + 1; r < 4; ++r) if (fabs(a[r][c]) > fabs(a[p][c])) p = r;
        if (p != c) for (int j = 0; j < 8; ++j) { double t = a[c][j]; a[c][j] = a[p][j]; a[p][j] = t; }
        double d = a[c][c];
        for (int j = 0; j < 8; ++j) a[c][j] /= d;
        for (int r = 0; r < 4; ++r) if (r != c) { double f = a[r][c]; for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j]; }
    }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) inv[4 * i + j] = a[i][4 + j];
}

int nerfb200_depth_type2(int H, int W, const double* K, const double* c2w, double scale_factor, const float* pred_depth,
                         float* out, void* stream) {
    NB_CHECK_ARG(H > 0 && W > 0 && K && c2w && pred_depth && out && scale_factor != 0.0, "depth_type2: bad arguments");
    CamD cam, inv;
    for (int i = 0; i < 9; ++i) cam.K[i] = inv.K[i] = K[i];
    for (int i = 0; i < 16; ++i) cam.c2w[i] = c2w[i];
    invert_rigid4(c2w, inv.c2w);
    int64_t n = (int64_t)H * W;
    depth_type2_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(W, n, cam, inv, 1.0 / scale_factor, pred_depth, out);
    NB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
