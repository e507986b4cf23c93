/*
 * nerfb200.h -- C ABI of the B200-native NeRF ray-march hot path.
 *
 * The reference (thatbrguy/nerf-tf2) has no native/FFI layer at all: its hot path is Python
 * over stock TensorFlow kernels. The entry points below are what a ctypes binding of that
 * path binds (INTEGRATION.md shows the reference-side stub); each one cites the reference
 * function (file:line under /root/reference) whose arithmetic it replaces.
 *
 * Conventions
 *   - every function returns 0 on success or a non-zero code (a cudaError_t value, or one of
 *     NERFB200_E*); nerfb200_last_error() returns a thread-local message for the last failure.
 *   - all pointers are DEVICE pointers owned by the caller (torch), fp32 row-major unless
 *     stated; the library owns only the opaque nerfb200_ctx (packed tensor-core weights).
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it.
 *   - no exceptions cross the ABI; no torch/C++ types in any signature.
 *   - B = rays, S = samples per ray, R = B*S network rows, row = ray*S + sample
 *     (utils/ray_utils.py:258,398,474).
 */
#ifndef NERFB200_H
#define NERFB200_H

#include <stdint.h>

#if defined(__GNUC__)
#define NERFB200_API __attribute__((visibility("default")))
#else
#define NERFB200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define NERFB200_ABI_VERSION 3

#define NERFB200_EINVAL   10001   /* bad argument                      */
#define NERFB200_ENOTSUP  10002   /* shape outside the supported range */
#define NERFB200_ESTATE   10003   /* e.g. weights not packed           */

/* MLP arithmetic selector (`precision` arguments) */
#define NERFB200_FP32 0   /* fp32 FMA on CUDA cores: the exact-arithmetic check path        */
#define NERFB200_BF16 1   /* bf16 operands, fp32 accumulate, tcgen05/TMEM fused kernel       */
#define NERFB200_FP16 2   /* fp16 operands, fp32 accumulate, tcgen05/TMEM fused kernel       */
#define NERFB200_TF32 3   /* tf32 operands (tcgen05 kind::tf32), fp32 accumulate: what TensorFlow's fp32 MatMul is on
                             Ampere-and-later GPUs (core/model.py:366-387 Dense layers); render path only */

#define NERFB200_COARSE 0
#define NERFB200_FINE   1

/* Parameters of ONE 8x256 model in the order of Keras' `model.trainable_variables` for the functional
 * model of core/model.py:334-394 -- layers sorted by decreasing depth from the outputs [rgb, sigma],
 * ties by output-first traversal -- i.e. dense_0..dense_9, rgb, sigma; each kernel [in,out] row-major
 * then its bias. Concatenating `model.get_weights()` in order gives exactly this block. */
#define NERFB200_PARAMS_PER_MODEL 595844
#define NERFB200_NUM_VARS_PER_MODEL 24
#define NERFB200_PARAMS_TOTAL (2 * NERFB200_PARAMS_PER_MODEL)  /* coarse then fine */

typedef struct nerfb200_ctx nerfb200_ctx;
typedef struct nerfb200_peer nerfb200_peer;   /* data-parallel gradient exchange over peer memory (below) */

NERFB200_API const char* nerfb200_last_error(void);
NERFB200_API int nerfb200_abi_version(void);
/* Number of kernels this library has launched in the process (bench.py reports the delta over the
 * timed region as `gpu_launches`). */
NERFB200_API int64_t nerfb200_launch_count(void);

/* Offsets (in floats) of the 24 variables of one model inside its flat parameter block,
 * order = model.trainable_variables (kernel, bias per layer). offsets[24] = total. */
NERFB200_API int nerfb200_param_offsets(int64_t* offsets /* [25] */);

/* ---- a1/a2: camera rays ---------------------------------------------------------------
 * get_rays (utils/ray_utils.py:6-51; fp64 maths, cast to fp32 as base_dataset.py:849-852) and
 * get_rays_tf (utils/ray_utils.py:53-106; fp32 maths). Ray id = row*W + col; writes rays
 * [ray0, ray0+n_rays) to rays_o/rays_d[n_rays,3]. K = 3x3 intrinsic, c2w = 4x4, row-major
 * HOST arrays (they are 25 scalars; passed by value to the kernel). */
NERFB200_API int nerfb200_get_rays(int H, int W, const double* K, const double* c2w, int64_t ray0,
                      int64_t n_rays, float* rays_o, float* rays_d, void* stream);
NERFB200_API int nerfb200_get_rays_f32(int H, int W, const float* K, const float* c2w, int64_t ray0,
                          int64_t n_rays, float* rays_o, float* rays_d, void* stream);
/* Sample-mode training input (core/base_dataset.py:555-621 computes all H*W rays then gathers):
 * rays for an explicit list of pixel ids (int32 device array), fp32 maths as get_rays_tf. */
NERFB200_API int nerfb200_get_rays_at(int H, int W, const float* K, const float* c2w, const int32_t* pixel_ids,
                         int64_t n_rays, float* rays_o, float* rays_d, void* stream);

/* ---- a3: stratified sampler ------------------------------------------------------------
 * create_input_batch_coarse_model (utils/ray_utils.py:137-274). near/far [B]. perturb!=0:
 * t = left + u*width with u = u_coarse[B,Nc] if non-NULL else Philox(seed, ray0+ray, sample);
 * perturb==0: bin mid-points (with the bin_widths fix of SURVEY.md App. B1).
 * Outputs t_vals[B,Nc], bin_edges[B,Nc+1] (left = [:, :-1], right = [:, 1:]). */
NERFB200_API int nerfb200_sample_coarse(int64_t B, int Nc, int lin_inv_depth, int perturb, const float* near,
                           const float* far, const float* u_coarse, uint64_t seed, const int64_t* step_state,
                           int64_t ray0, float* t_vals, float* bin_edges, void* stream);
/* `step_state` (may be NULL) -- device-resident step state int64[2] = {optimizer iterations, sampling step}: when given,
 * the samplers use seed ^ step_state[1] and nerfb200_adam_step reads the iteration count from step_state[0], so that a
 * whole training step captured in a CUDA graph advances its noise and its learning-rate schedule on replay
 * (nerfb200_step_advance increments both counters on the stream). */

/* xyz/dir network inputs exactly as the reference materialises them (utils/ray_utils.py:251-258):
 * xyz = o + t*d (separate multiply and add), dirs broadcast. For parity tests and the FP32
 * path; the tensor-core kernel fuses this and never writes xyz to HBM. */
NERFB200_API int nerfb200_make_inputs(int64_t B, int S, const float* rays_o, const float* rays_d,
                         const float* t_vals, float* xyz /* [B*S,3] */, float* dirs /* [B*S,3] */,
                         void* stream);

/* ---- a4: positional encoding (core/model.py:289-332) -----------------------------------
 * out[R, 3+6L]: [x, then for d in 0..2, l in 0..L-1: sin(x_d*m_l), cos(x_d*m_l)],
 * m_l = fl32(2^l)*fl32(pi), one fp32 multiply then accurate sinf/cosf. */
NERFB200_API int nerfb200_positional_encode(int64_t R, int L, const float* x /* [R,3] */, float* out, void* stream);

/* ---- context: packed weights ------------------------------------------------------------ */
NERFB200_API int nerfb200_create(nerfb200_ctx** out);
NERFB200_API int nerfb200_destroy(nerfb200_ctx* ctx);
/* (Re)pack the fp32 master parameters (flat [NERFB200_PARAMS_TOTAL], coarse then fine) into the
 * tensor-core operand image (K-major 128B-swizzled UMMA tiles, bf16 and fp16 copies).
 * Must be called after every optimiser step before the next forward. */
NERFB200_API int nerfb200_pack_weights(nerfb200_ctx* ctx, const float* flat_params, void* stream);
/* Context options.
 *   NERFB200_OPT_PRECISE_LAST (default 1; 0 off; 2 = training forwards, i.e. stash != NULL, too): tensor-core forwards
 *     recompute sigma of the LAST sample of every
 *     ray (row ray*S + S-1, S >= 2) with error-compensated split operands, because the reference's
 *     delta_last = 1e10 (utils/ray_utils.py:459-468) turns a rounding-induced sign flip of that one ReLU
 *     output into a jump of alpha_last from 0 to 1.
 *   NERFB200_OPT_PACK_MASK (default 7): precisions pack_weights produces images for, bit 0 bf16, bit 1 fp16,
 *     bit 2 tf32 (a training loop packs after every step and needs only its own precision).
 *   NERFB200_OPT_DEBUG (default 0): developer cycle counters of the fused forward (tools/tc_debug.py).
 * PRECISE_LAST and PACK_MASK take effect at the next nerfb200_pack_weights. */
#define NERFB200_OPT_PRECISE_LAST 1
#define NERFB200_OPT_PACK_MASK    2
#define NERFB200_OPT_DEBUG        3
NERFB200_API int nerfb200_set_option(nerfb200_ctx* ctx, int option, int value);

/* ---- a4+a5: fused encoding + 8x256 MLP forward (core/model.py:334-394) -------------------
 * rows are generated on the fly from rays: xyz = o + t*d, dir = d. Outputs rgb[R,3] (sigmoid)
 * and sigma[R] (relu). `flat_params` is the fp32 master block (used by NERFB200_FP32 and for
 * the fp32 sigma/rgb heads); tensor-core precisions read the image from pack_weights.
 * workspace: device scratch of at least nerfb200_mlp_workspace_bytes(...) bytes (may be NULL
 * when that is 0). If `stash` is non-NULL (training), activations needed by
 * nerfb200_mlp_backward are written there (nerfb200_mlp_stash_bytes). */
NERFB200_API int64_t nerfb200_mlp_workspace_bytes(int64_t R, int precision, int training);
NERFB200_API int64_t nerfb200_mlp_stash_bytes(int64_t R, int precision);
NERFB200_API int nerfb200_mlp_forward(nerfb200_ctx* ctx, int which, int64_t B, int S, const float* rays_o,
                         const float* rays_d, const float* t_vals, const float* flat_params,
                         float* rgb, float* sigma, int precision, void* workspace, void* stash,
                         void* stream);

/* ---- a6: MLP backward (tf.GradientTape over core/model.py:148-170) -----------------------
 * d_rgb[R,3], d_sigma[R] are dLoss/d(outputs). Accumulates (+=) into flat_grads (same layout as
 * flat_params; caller zeroes it once per step). No dX for the inputs: sample positions are
 * constants for autodiff (stop_gradient, utils/ray_utils.py:377). */
NERFB200_API int nerfb200_mlp_backward(nerfb200_ctx* ctx, int which, int64_t B, int S, const float* rays_o,
                          const float* rays_d, const float* t_vals, const float* flat_params,
                          const float* d_rgb, const float* d_sigma, float* flat_grads,
                          int precision, void* workspace, void* stash, void* stream);
/* The same backward pass as two separately launchable phases (tensor-core precisions only; FP32 returns
 * NERFB200_ENOTSUP): _data writes the per-layer gradient stash into `workspace` (HBM-write bound), _weights
 * reads both stashes and accumulates into flat_grads (HBM-read bound). max_sms > 0 caps the SMs a phase
 * occupies (0 = all), so a caller can run one model's _weights next to the other model's _data on disjoint
 * SMs and two streams; each model then needs its own workspace. mlp_backward == _data then _weights, max_sms 0. */
NERFB200_API int nerfb200_mlp_backward_data(nerfb200_ctx* ctx, int which, int64_t B, int S, const float* flat_params,
                               const float* d_rgb, const float* d_sigma, int precision, void* workspace,
                               void* stash, int max_sms, void* stream);
NERFB200_API int nerfb200_mlp_backward_weights(nerfb200_ctx* ctx, int which, int64_t B, int S, float* flat_grads,
                                  int precision, void* workspace, void* stash, int max_sms, void* stream);

/* ---- a11: NeRF.forward (core/model.py:57-125) as ONE call --------------------------------------------
 * stratified sampling -> coarse MLP -> integrator -> hierarchical sampling -> fine MLP -> integrator for B rays, every
 * launch on `stream`, the intermediates (sample positions, per-sample rgb/sigma, coarse weights) in `workspace`
 * (nerfb200_forward_workspace_bytes): nothing is allocated and nothing returns to the host in between. Tensor-core
 * precisions only (bf16 / fp16 / tf32; pack_weights first). u_coarse / u_fine / seed / step_state / ray0 as in the
 * samplers. Outputs per model: pred_rgb [B,3], pred_depth [B], acc_map [B], weights [B,S] (the weights may be NULL) --
 * the four entries of the reference's result dictionaries (utils/ray_utils.py:546-551). */
NERFB200_API int64_t nerfb200_forward_workspace_bytes(int64_t B, int Nc, int Nf);
NERFB200_API int nerfb200_forward(nerfb200_ctx* ctx, int64_t B, int Nc, int Nf, int lin_inv_depth, int perturb, int white_bg,
                     const float* rays_o, const float* rays_d, const float* near, const float* far,
                     const float* u_coarse, const float* u_fine, uint64_t seed, const int64_t* step_state, int64_t ray0,
                     const float* flat_params, int precision, void* workspace,
                     float* coarse_rgb, float* coarse_depth, float* coarse_acc, float* coarse_weights,
                     float* fine_rgb, float* fine_depth, float* fine_acc, float* fine_weights, void* stream);

/* ---- a7-a9: volume-rendering integrator --------------------------------------------------
 * sigma_to_alpha / compute_weights / post_process_model_output (utils/ray_utils.py:408-551):
 * delta_i = t_{i+1}-t_i, delta_last = 1e10; alpha = 1-exp(-sigma*delta);
 * w = alpha * cumprod_exclusive(1-alpha+1e-10); rgb/depth/acc = sum(w*.); white_bg: rgb += 1-acc.
 * weights may be NULL (skips the [B,S] store). Warp-per-ray shuffle scan; 2 <= S <= 1024. */
NERFB200_API int nerfb200_composite_fwd(int64_t B, int S, const float* sigma, const float* rgb,
                           const float* t_vals, int white_bg, float* weights, float* pred_rgb,
                           float* pred_depth, float* acc_map, void* stream);
/* Backward of the above w.r.t. sigma and rgb given d_pred_rgb[B,3] (pred_depth/acc are not in
 * the loss; acc enters through white_bg). Recomputes alpha/T from sigma,t. */
NERFB200_API int nerfb200_composite_bwd(int64_t B, int S, const float* sigma, const float* rgb,
                           const float* t_vals, int white_bg, const float* d_pred_rgb,
                           float* d_sigma, float* d_rgb, void* stream);

/* Training form (NeRF.train_step, core/model.py:148-170): nerfb200_composite_fwd, the loss of
 * nerfb200_mse_loss_grad on this output (loss += mean((pred-gt)^2) with the mean over B_global*3 values; metric[0] +=
 * sum((gt-pred)^2), metric[1] += B when metric != NULL) and nerfb200_composite_bwd with d_pred = 2*(pred-gt)/(B_global*3),
 * as ONE launch over values that stay in registers. Same arithmetic as the three separate calls (the loss is summed per
 * ray instead of per 256 values). weights and metric may be NULL. */
NERFB200_API int nerfb200_composite_train(int64_t B, int S, const float* sigma, const float* rgb, const float* t_vals,
                           int white_bg, const float* rgb_gt, int64_t B_global, float* weights, float* pred_rgb,
                           float* pred_depth, float* acc_map, float* d_sigma, float* d_rgb, float* loss, float* metric,
                           void* stream);

/* ---- a10: hierarchical (inverse-transform) sampler ----------------------------------------
 * create_input_batch_fine_model (utils/ray_utils.py:276-406): pdf/cdf from (w+1e-5), upper-bound
 * searchsorted over the Nc-1 inner CDF edges, inversion with the pdf<1e-8 mask, then
 * sort(concat(t_coarse, t_fine)). u = u_fine[B,Nf] if non-NULL (reproduces a given
 * tf.random.uniform draw: searchsorted indices bit-exact on the kernel's fp32 CDF, output the
 * exact ascending sort). With u_fine NULL the uniforms come from Philox(seed, ray0+ray): the
 * 64/128 and 128/256 shapes draw the ORDER STATISTICS of Nf i.i.d. U[0,1) directly (normalised
 * partial sums of Nf+1 exponentials - the same joint distribution as drawing Nf uniforms and
 * sorting, which is what the reference does at :355,:385), other shapes draw u_j i.i.d.; both
 * are pure functions of (seed, global ray id), hence independent of sharding and chunking.
 * Optional debug outputs (NULL to skip): piece_idxs[B,Nf] int32, cdf[B,Nc+1] (the fp32 CDF the
 * indices were searched in), t_fine[B,Nf] in the order of u. Nc in {32..256 step 32}, Nf in
 * {32,64,128,256,512}. */
NERFB200_API int nerfb200_sample_fine(int64_t B, int Nc, int Nf, const float* bin_weights,
                         const float* bin_edges, const float* t_coarse, const float* u_fine,
                         uint64_t seed, const int64_t* step_state, int64_t ray0, float* t_sorted /* [B,Nc+Nf] */,
                         int32_t* piece_idxs, float* cdf, float* t_fine, void* stream);

/* ---- a12: loss + optimiser -----------------------------------------------------------------
 * Keras MeanSquaredError over [B,3] (core/model.py:157-168): adds mean((pred-gt)^2) to *loss and
 * writes d_pred = 2*(pred-gt)/(B_global*3). Also accumulates the PSNRMetric state
 * (core/ops.py:204-220): metric[0] += sum((gt-pred)^2), metric[1] += B, when metric != NULL. */
NERFB200_API int nerfb200_mse_loss_grad(int64_t B, int64_t B_global, const float* pred_rgb, const float* rgb_gt,
                           float* d_pred, float* loss, float* metric, void* stream);
/* Keras Adam (OptimizerV2, beta 0.9/0.999, eps 1e-7) with ExponentialDecay(5e-4, 500000, 0.1)
 * (core/model.py:413-418); `iterations` is the counter BEFORE the step. Fused over the flat
 * parameter block. */
NERFB200_API int nerfb200_adam_step(int64_t n, float* params, const float* grads, float* m, float* v,
                       int64_t iterations, const int64_t* step_state /* overrides `iterations` when non-NULL */, void* stream);
NERFB200_API int nerfb200_step_advance(int64_t* step_state /* device int64[2] */, void* stream);

/* ---- (e) data-parallel training: the gradient exchange over NVLink peer memory ------------------
 * The reference applies the gradients of ONE device's tape (core/model.py:148-171); data parallel, the flat
 * buffer [coarse gradient | fine gradient | loss,0,0,0] has to be summed over the ranks before the replicated
 * Adam. One process per GPU of one NVSwitch box (world <= 8):
 *   peer_create   allocates this rank's block (a small flag header + n_floats floats, zeroed) on the current device;
 *   peer_buffer   the n_floats floats -- the caller accumulates its local gradient there;
 *   peer_handle   64 opaque bytes (a CUDA IPC handle) to be exchanged between the ranks by any means
 *                 (torch.distributed.all_gather here);
 *   peer_connect  maps every other rank's block (`handles` = world x 64 bytes, in rank order);
 *   peer_allreduce  ONE kernel: flag barrier, every rank sums its 1/world slice of all ranks' buffers over
 *                 NVLink in rank order and stores the sum into that slice of EVERY rank's buffer, flag barrier.
 *                 The launch completes when the local buffer holds the full sum; all ranks receive bit-identical
 *                 sums. Every rank must launch it the same number of times (like a collective). A rank whose peers
 *                 do not arrive within the timeout (default 120 s) traps instead of hanging;
 *   peer_allreduce_adam  the same exchange followed, in the same launch, by nerfb200_adam_step over the first n
 *                 floats of the summed buffer (params/m/v: local, 16-byte aligned, n a multiple of 4): identical
 *                 arithmetic to peer_allreduce + adam_step.
 *   peer_attach   instead of create/handle/connect, for a caller that has ALREADY mapped the ranks' blocks (here:
 *                 torch symmetric memory): `blocks` = world device pointers, the block of every rank as mapped in
 *                 this process, each 4096 header bytes (zeroed by the caller) + n_floats floats; `multicast_block` =
 *                 the same block through an NVSwitch multicast (NVLS) mapping, or NULL. With a multicast mapping
 *                 the slice is summed by the switch (multimem.ld_reduce) and replicated by the switch (multimem.st):
 *                 1/world of the NVLink traffic; the order of that sum is the switch's, identical on all ranks.
 * Flags carry an epoch kept in device memory, so both launches can be captured in a CUDA graph and replayed. */
NERFB200_API int nerfb200_peer_attach(int world, int rank, int64_t n_floats, void* const* blocks, void* multicast_block,
                                      nerfb200_peer** peer);
NERFB200_API int nerfb200_peer_create(int world, int rank, int64_t n_floats, nerfb200_peer** peer);
NERFB200_API int nerfb200_peer_buffer(nerfb200_peer* peer, float** buffer);
NERFB200_API int nerfb200_peer_handle(nerfb200_peer* peer, unsigned char* handle64);
NERFB200_API int nerfb200_peer_connect(nerfb200_peer* peer, const unsigned char* handles);
NERFB200_API int nerfb200_peer_set_timeout(nerfb200_peer* peer, int seconds);
NERFB200_API int nerfb200_peer_allreduce(nerfb200_peer* peer, void* stream);
NERFB200_API int nerfb200_peer_allreduce_adam(nerfb200_peer* peer, int64_t n, float* params, float* m, float* v,
                                 int64_t iterations, const int64_t* step_state, void* stream);
NERFB200_API int nerfb200_peer_status(nerfb200_peer* peer, int* status /* 0, or the barrier (1/2) a wait timed out at */);
/* diagnostic: %globaltimer (ns) of this rank's latest exchange -- start, barrier A passed, own slice done (first CTA);
 * all CTAs done, barrier B passed, end (last CTA); [6] = the first CTA's loads returned and stores issued, before its
 * system fence. Synchronous copy: call it on an idle stream. */
NERFB200_API int nerfb200_peer_profile(nerfb200_peer* peer, unsigned long long* ns7);
NERFB200_API int nerfb200_peer_disconnect(nerfb200_peer* peer);   /* unmaps the other ranks' blocks */
NERFB200_API int nerfb200_peer_destroy(nerfb200_peer* peer);      /* + frees this rank's block: only after EVERY rank has disconnected */

/* ---- a15: depth map type_2 (utils/ray_utils.py:122-130) -------------------------------------
 * z of the point o + d*depth/scale in the camera frame. */
NERFB200_API int nerfb200_depth_type2(int H, int W, const double* K, const double* c2w, double scale_factor,
                         const float* pred_depth, float* out, void* stream);

/* ---- callers either side of the path -----------------------------------
 * next-2, sample-mode training input (core/base_dataset.py:555-621): B uniform pixel ids in
 * [0, n_pixels) (tf.random.uniform(..., dtype=int32) -> Philox(seed, step)); rays of those pixels come
 * from nerfb200_get_rays_at; colours are gathered from a device-resident uint8 [H*W,3] image as
 * float(img)/255 (:596-598). */
NERFB200_API int nerfb200_sample_pixels(int64_t B, int64_t n_pixels, uint64_t seed, uint64_t step,
                                        int32_t* pixel_ids, void* stream);
NERFB200_API int nerfb200_gather_rgb_u8(int64_t B, const uint8_t* image, const int32_t* pixel_ids,
                                        float* rgb, void* stream);
/* next-3, per-view post-processing (main/eval.py:53-64, main/render.py:96-100): out_u8 =
 * uint8(clip(pred*255,0,255)) and, if gt_u8 is given, *sq_err += sum((gt/255 - clip(pred*255,0,255)/255)^2)
 * over the n_values = H*W*3 channel values (psnr_metric_numpy = -10*log10(sq_err / n_values)). */
NERFB200_API int nerfb200_postprocess_rgb(int64_t n_values, const float* pred_rgb, const uint8_t* gt_u8,
                                          uint8_t* out_u8, double* sq_err, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NERFB200_H */
