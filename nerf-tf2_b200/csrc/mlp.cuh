// Internal interface between the MLP translation units.
#pragma once
#include "common.cuh"

struct nerfb200_ctx {
    int device = 0;
    int num_sms = 148;
    // tensor-core operand images, one per (precision, model): see mlp_tc.cu for the layout
    void* packed[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};      // [bf16|fp16][coarse|fine]
    void* packed_lo[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // W - fl16(W): the split launch of the last-sample rows
    void* packed_bwd[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // W^T images for backward-data
    void* packed_tf32[2] = {nullptr, nullptr};                          // tf32 image per model (mlp_tc_tf32.cu)
    float* head_params[2] = {nullptr, nullptr};                         // fp32 biases + sigma head per model
    bool packed_valid = false;
    int packed_mask = 0;            // precisions the current images were packed for
    bool packed_precise = false;    // ... and whether the lo images were packed with them
    // options (nerfb200_set_option)
    int precise_last = 1;           // NERFB200_OPT_PRECISE_LAST
    int pack_mask = 7;              // NERFB200_OPT_PACK_MASK: bit 0 bf16, bit 1 fp16, bit 2 tf32
    int debug = 0;                  // NERFB200_OPT_DEBUG: cycle-counter instantiation of the pair kernel
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
int check_device(const nerfb200_ctx* ctx, const char* fn);
// mlp_tc_tf32.cu: the kind::tf32 render kernel (inference only)
int tf32_create(nerfb200_ctx* ctx);
void tf32_destroy(nerfb200_ctx* ctx);
int tf32_pack(nerfb200_ctx* ctx, const float* flat_params, cudaStream_t st);
int tf32_forward(nerfb200_ctx* ctx, int which, int64_t B, int S, const float* ro, const float* rd, const float* t,
                 float* rgb, float* sigma, cudaStream_t st);
// sigma of the last sample of every ray with split (hi + lo) 16-bit operands; no-op when the option is off or S < 2
int tc_precise_last(nerfb200_ctx* ctx, int which, int half, int64_t B, int S, const float* ro, const float* rd, const float* t,
                    float* sigma, void* stash, cudaStream_t st);
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
