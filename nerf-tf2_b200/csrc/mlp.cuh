// Internal interface between the MLP translation units.
#pragma once
#include "common.cuh"

struct nerfb200_ctx {
    int device = 0;
    int num_sms = 148;
    // tensor-core operand images, one per (precision, model): see mlp_tc.cu for the layout
    void* packed[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [bf16|fp16][coarse|fine]
    void* packed_bwd[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // W^T images for backward-data
    float* head_params[2] = {nullptr, nullptr};                       // fp32 biases + sigma head per model
    bool packed_valid = false;
    int replicas = 1;                                                 // identical copies of each packed image
};

namespace nb {

constexpr int64_t kRefChunkRows = 65536;

// mlp_ref.cu
int64_t ref_workspace_bytes(int64_t R, int training);
int64_t ref_stash_bytes(int64_t R);
int ref_forward(cudaStream_t st, const float* P, int64_t B, int S, const float* ro, const float* rd, const float* t,
                float* rgb, float* sigma, void* workspace, void* stash);
int ref_backward(cudaStream_t st, const float* P, int64_t B, int S, const float* d_rgb, const float* d_sigma,
                 float* G, void* workspace, void* stash);

// mlp_tc.cu / mlp_tc_train.cu
int tc_train_create(nerfb200_ctx* ctx);
void tc_train_destroy(nerfb200_ctx* ctx);
int tc_train_pack(nerfb200_ctx* ctx, const float* flat_params, cudaStream_t st);
int tc_create(nerfb200_ctx* ctx);
void tc_destroy(nerfb200_ctx* ctx);
int tc_pack_weights(nerfb200_ctx* ctx, const float* flat_params, cudaStream_t st);
int64_t tc_workspace_bytes(int64_t R, int training);
int64_t tc_stash_bytes(int64_t R);
int tc_forward(nerfb200_ctx* ctx, int which, int half, int64_t B, int S, const float* ro, const float* rd,
               const float* t, float* rgb, float* sigma, void* workspace, void* stash, cudaStream_t st);
int tc_backward(nerfb200_ctx* ctx, int which, int half, int64_t B, int S, const float* ro, const float* rd,
                const float* t, const float* flat_params, const float* d_rgb, const float* d_sigma, float* flat_grads,
                void* workspace, void* stash, cudaStream_t st);
int tc_backward_data(nerfb200_ctx* ctx, int which, int half, int64_t B, int S, const float* flat_params, const float* d_rgb,
                     const float* d_sigma, void* workspace, void* stash, int max_sms, cudaStream_t st);
int tc_backward_weights(nerfb200_ctx* ctx, int which, int half, int64_t B, int S, float* flat_grads, void* workspace,
                        void* stash, int max_sms, cudaStream_t st);

}  // namespace nb
