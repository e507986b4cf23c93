// Loss, metric state and the fused Adam step of NeRF.train_step (core/model.py:127-180, :413-418).
#include "common.cuh"

#include <math.h>

namespace nb {

// Keras MeanSquaredError over [B,3]: loss += sum((pred-gt)^2)/(B_global*3); d_pred = 2*(pred-gt)/(B_global*3).
// PSNRMetric.update_state (core/ops.py:204-220): metric[0] += sum((gt-pred)^2), metric[1] += B.
__global__ void mse_loss_grad_kernel(int64_t B, float inv_n, const float* __restrict__ pred, const float* __restrict__ gt,
                                     float* __restrict__ d_pred, float* __restrict__ loss, float* __restrict__ metric) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float sq = 0.f;
    if (i < B * 3) {
        float d = pred[i] - gt[i];
        if (d_pred) d_pred[i] = 2.f * d * inv_n;
        sq = d * d;
    }
    sq = warp_sum(sq);
    __shared__ float part[8];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) part[w] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += part[k];
        if (loss) atomicAdd(loss, s * inv_n);
        if (metric) {
            atomicAdd(metric, s);
            if (blockIdx.x == 0) atomicAdd(metric + 1, (float)B);
        }
    }
}

__global__ void adam_kernel(int64_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, float lr_t, const int64_t* __restrict__ step_dev) {
    if (step_dev) {          // device-resident iteration counter (CUDA-graph replays): same formula, evaluated once per block
        __shared__ float s_lr_t;
        if (threadIdx.x == 0) s_lr_t = adam_lr_t(step_dev[0]);
        __syncthreads();
        lr_t = s_lr_t;
    }
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float pi = p[i], mi = m[i], vi = v[i];
    adam_update(pi, mi, vi, g[i], lr_t);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
}

__global__ void step_advance_kernel(int64_t* __restrict__ step) { step[0] += 1; step[1] += 1; }

}  // namespace nb

using namespace nb;

extern "C" {

int nerfb200_mse_loss_grad(int64_t B, int64_t B_global, const float* pred_rgb, const float* rgb_gt, float* d_pred,
                           float* loss, float* metric, void* stream) {
    NB_CHECK_ARG(B >= 0 && B_global >= B && pred_rgb && rgb_gt, "mse_loss_grad: bad arguments");
    if (B == 0) return 0;
    float inv_n = 1.0f / (float)(B_global * 3);
    mse_loss_grad_kernel<<<(unsigned)((B * 3 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(B, inv_n, pred_rgb, rgb_gt,
                                                                                          d_pred, loss, metric);
    NB_LAUNCH_CHECK();
    return 0;
}

int nerfb200_adam_step(int64_t n, float* params, const float* grads, float* m, float* v, int64_t iterations,
                       const int64_t* step_state, void* stream) {
    NB_CHECK_ARG(n >= 0 && params && grads && m && v && iterations >= 0, "adam_step: bad arguments");
    if (n == 0) return 0;
    const float lr_t = adam_lr_t(iterations);
    adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, params, grads, m, v, lr_t, step_state);
    NB_LAUNCH_CHECK();
    return 0;
}

int nerfb200_step_advance(int64_t* step_state, void* stream) {
    NB_CHECK_ARG(step_state != nullptr, "step_advance: NULL state");
    step_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_state);
    NB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
