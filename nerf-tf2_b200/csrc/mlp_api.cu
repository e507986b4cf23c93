// C-ABI entry points of the MLP: context, weight packing, forward/backward dispatch by precision.
#include "common.cuh"
#include "mlp.cuh"

using namespace nb;

extern "C" {

int nerfb200_create(nerfb200_ctx** out) {
    NB_CHECK_ARG(out != nullptr, "create: NULL output");
    nerfb200_ctx* ctx = new nerfb200_ctx();
    int rc = tc_create(ctx);
    if (rc) { tc_destroy(ctx); delete ctx; return rc; }
    *out = ctx;
    return 0;
}

int nerfb200_destroy(nerfb200_ctx* ctx) {
    if (!ctx) return 0;
    tc_destroy(ctx);
    delete ctx;
    return 0;
}

int nerfb200_pack_weights(nerfb200_ctx* ctx, const float* flat_params, void* stream) {
    NB_CHECK_ARG(ctx && flat_params, "pack_weights: NULL argument");
    return tc_pack_weights(ctx, flat_params, (cudaStream_t)stream);
}

int nerfb200_set_option(nerfb200_ctx* ctx, int option, int value) {
    NB_CHECK_ARG(ctx != nullptr, "set_option: NULL context");
    switch (option) {
        case NERFB200_OPT_PRECISE_LAST:
            NB_CHECK_ARG(value >= 0 && value <= 2, "set_option(PRECISE_LAST): 0 off, 1 render forwards, 2 training forwards too");
            ctx->precise_last = value;
            return 0;
        case NERFB200_OPT_PACK_MASK:
            NB_CHECK_ARG(value >= 1 && value <= 7, "set_option(PACK_MASK): value must be a non-empty subset of bits 0..2");
            ctx->pack_mask = value;
            return 0;
        case NERFB200_OPT_DEBUG: ctx->debug = value; return 0;
        default: NB_CHECK_ARG(false, "set_option: unknown option %d", option);
    }
    return 0;
}

int64_t nerfb200_mlp_workspace_bytes(int64_t R, int precision, int training) {
    if (R < 0) return -1;
    if (precision == NERFB200_TF32) return 0;
    return precision == NERFB200_FP32 ? ref_workspace_bytes(R, training) : tc_workspace_bytes(R, training);
}

int64_t nerfb200_mlp_stash_bytes(int64_t R, int precision) {
    if (R < 0) return -1;
    if (precision == NERFB200_TF32) return 0;
    return precision == NERFB200_FP32 ? ref_stash_bytes(R) : tc_stash_bytes(R);
}

static int check_common(const char* fn, nerfb200_ctx* ctx, int which, int64_t B, int S, int precision) {
    NB_CHECK_ARG(ctx != nullptr, "%s: NULL context", fn);
    NB_CHECK_ARG(which == NERFB200_COARSE || which == NERFB200_FINE, "%s: which must be 0 (coarse) or 1 (fine)", fn);
    NB_CHECK_ARG(B >= 0 && S >= 1, "%s: bad shape B=%lld S=%d", fn, (long long)B, S);
    NB_CHECK_ARG(precision >= NERFB200_FP32 && precision <= NERFB200_TF32, "%s: unknown precision %d", fn, precision);
    return 0;
}

int nerfb200_mlp_forward(nerfb200_ctx* ctx, int which, int64_t B, int S, const float* rays_o, const float* rays_d,
                         const float* t_vals, const float* flat_params, float* rgb, float* sigma, int precision,
                         void* workspace, void* stash, void* stream) {
    int rc = check_common("mlp_forward", ctx, which, B, S, precision);
    if (rc) return rc;
    NB_CHECK_ARG(rays_o && rays_d && t_vals && rgb && sigma, "mlp_forward: NULL tensor");
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == NERFB200_FP32) {
        NB_CHECK_ARG(flat_params != nullptr, "mlp_forward(fp32): flat_params required");
        return ref_forward(st, flat_params + (int64_t)which * kParamsPerModel, B, S, rays_o, rays_d, t_vals, rgb, sigma,
                           workspace, stash);
    }
    if (precision == NERFB200_TF32) {
        if (stash) { set_error("mlp_forward: tf32 is a render precision; training uses bf16, fp16 or fp32"); return NERFB200_ENOTSUP; }
        return tf32_forward(ctx, which, B, S, rays_o, rays_d, t_vals, rgb, sigma, st);
    }
    return tc_forward(ctx, which, precision == NERFB200_FP16, B, S, rays_o, rays_d, t_vals, rgb, sigma, workspace, stash, st);
}

int nerfb200_mlp_backward(nerfb200_ctx* ctx, int which, int64_t B, int S, const float* rays_o, const float* rays_d,
                          const float* t_vals, const float* flat_params, const float* d_rgb, const float* d_sigma,
                          float* flat_grads, int precision, void* workspace, void* stash, void* stream) {
    int rc = check_common("mlp_backward", ctx, which, B, S, precision);
    if (rc) return rc;
    NB_CHECK_ARG(flat_params && d_rgb && d_sigma && flat_grads, "mlp_backward: NULL tensor");
    if (precision == NERFB200_TF32) { set_error("mlp_backward: tf32 is a render precision; training uses bf16, fp16 or fp32"); return NERFB200_ENOTSUP; }
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == NERFB200_FP32)
        return ref_backward(st, flat_params + (int64_t)which * kParamsPerModel, B, S, d_rgb, d_sigma,
                            flat_grads + (int64_t)which * kParamsPerModel, workspace, stash);
    return tc_backward(ctx, which, precision == NERFB200_FP16, B, S, rays_o, rays_d, t_vals, flat_params, d_rgb, d_sigma,
                       flat_grads, workspace, stash, st);
}

int nerfb200_mlp_backward_data(nerfb200_ctx* ctx, int which, int64_t B, int S, const float* flat_params, const float* d_rgb,
                               const float* d_sigma, int precision, void* workspace, void* stash, int max_sms, void* stream) {
    int rc = check_common("mlp_backward_data", ctx, which, B, S, precision);
    if (rc) return rc;
    if (precision == NERFB200_FP32 || precision == NERFB200_TF32) {
        set_error("mlp_backward_data: the phase-split backward exists for the 16-bit tensor-core precisions only");
        return NERFB200_ENOTSUP;
    }
    NB_CHECK_ARG(flat_params && d_rgb && d_sigma && max_sms >= 0, "mlp_backward_data: bad argument");
    if (B == 0) return 0;
    return tc_backward_data(ctx, which, precision == NERFB200_FP16, B, S, flat_params, d_rgb, d_sigma, workspace, stash, max_sms,
                            (cudaStream_t)stream);
}

int nerfb200_mlp_backward_weights(nerfb200_ctx* ctx, int which, int64_t B, int S, float* flat_grads, int precision,
                                  void* workspace, void* stash, int max_sms, void* stream) {
    int rc = check_common("mlp_backward_weights", ctx, which, B, S, precision);
    if (rc) return rc;
    if (precision == NERFB200_FP32 || precision == NERFB200_TF32) {
        set_error("mlp_backward_weights: the phase-split backward exists for the 16-bit tensor-core precisions only");
        return NERFB200_ENOTSUP;
    }
    NB_CHECK_ARG(flat_grads && max_sms >= 0, "mlp_backward_weights: bad argument");
    if (B == 0) return 0;
    return tc_backward_weights(ctx, which, precision == NERFB200_FP16, B, S, flat_grads, workspace, stash, max_sms,
                               (cudaStream_t)stream);
}

// ---- a11: the whole ray march of NeRF.forward (core/model.py:57-125) as ONE call -------------------------------------
// stratified sampling -> coarse MLP -> integrator -> hierarchical sampling -> fine MLP -> integrator, every launch on
// `stream`, intermediates in the caller's workspace (no allocation, nothing returns to the host in between).
static inline int64_t align256(int64_t n) { return (n + 255) & ~(int64_t)255; }

int64_t nerfb200_forward_workspace_bytes(int64_t B, int Nc, int Nf) {
    if (B < 0 || Nc < 2 || Nf < 1) return -1;
    const int64_t S = Nc + Nf;
    // t_c [B,Nc], edges [B,Nc+1], rgb_c [B*Nc,3], sigma_c [B*Nc], weights_c [B,Nc] (when the caller does not want them),
    // t_f [B,S], rgb_f [B*S,3], sigma_f [B*S]
    return align256(4 * B * Nc) + align256(4 * B * (Nc + 1)) + align256(12 * B * Nc) + align256(4 * B * Nc) + align256(4 * B * Nc) +
           align256(4 * B * S) + align256(12 * B * S) + align256(4 * B * S);
}

int nerfb200_forward(nerfb200_ctx* ctx, int64_t B, int Nc, int Nf, int lin_inv_depth, int perturb, int white_bg,
                     const float* rays_o, const float* rays_d, const float* near, const float* far, const float* u_coarse,
                     const float* u_fine, uint64_t seed, const int64_t* step_state, int64_t ray0, const float* flat_params,
                     int precision, void* workspace, float* c_rgb, float* c_depth, float* c_acc, float* c_weights, float* f_rgb,
                     float* f_depth, float* f_acc, float* f_weights, void* stream) {
    NB_CHECK_ARG(ctx != nullptr, "forward: NULL context");
    NB_CHECK_ARG(B >= 0 && Nc >= 2 && Nf >= 1, "forward: bad shape B=%lld Nc=%d Nf=%d", (long long)B, Nc, Nf);
    NB_CHECK_ARG(precision >= NERFB200_FP32 && precision <= NERFB200_TF32, "forward: unknown precision %d", precision);
    if (B == 0) return 0;
    NB_CHECK_ARG(rays_o && rays_d && near && far && workspace && c_rgb && c_depth && c_acc && f_rgb && f_depth && f_acc,
                 "forward: NULL tensor");
    if (precision == NERFB200_FP32) {
        set_error("forward: the one-call ray march exists for the tensor-core precisions (the fp32 check path needs its own MLP workspace)");
        return NERFB200_ENOTSUP;
    }
    const int S = Nc + Nf;
    uint8_t* w = (uint8_t*)workspace;
    auto take = [&](int64_t bytes) { float* p = (float*)w; w += align256(bytes); return p; };
    float* t_c = take(4 * B * Nc);
    float* edges = take(4 * B * (Nc + 1));
    float* rgb_c = take(12 * B * Nc);
    float* sig_c = take(4 * B * Nc);
    float* w_c = take(4 * B * Nc);
    float* t_f = take(4 * B * S);
    float* rgb_f = take(12 * B * S);
    float* sig_f = take(4 * B * S);
    if (c_weights) w_c = c_weights;
    int rc;
    if ((rc = nerfb200_sample_coarse(B, Nc, lin_inv_depth, perturb, near, far, u_coarse, seed, step_state, ray0, t_c, edges, stream))) return rc;
    if ((rc = nerfb200_mlp_forward(ctx, NERFB200_COARSE, B, Nc, rays_o, rays_d, t_c, flat_params, rgb_c, sig_c, precision, nullptr, nullptr, stream))) return rc;
    if ((rc = nerfb200_composite_fwd(B, Nc, sig_c, rgb_c, t_c, white_bg, w_c, c_rgb, c_depth, c_acc, stream))) return rc;
    if ((rc = nerfb200_sample_fine(B, Nc, Nf, w_c, edges, t_c, u_fine, seed, step_state, ray0, t_f, nullptr, nullptr, nullptr, stream))) return rc;
    if ((rc = nerfb200_mlp_forward(ctx, NERFB200_FINE, B, S, rays_o, rays_d, t_f, flat_params, rgb_f, sig_f, precision, nullptr, nullptr, stream))) return rc;
    return nerfb200_composite_fwd(B, S, sig_f, rgb_f, t_f, white_bg, f_weights, f_rgb, f_depth, f_acc, stream);
}

}  // extern "C"
