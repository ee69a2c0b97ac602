// host-side internals shared by the translation units of libsplendor_b200.so (not part of the ABI)
#pragma once
#include <cuda_runtime.h>
#include "../../include/splendor_b200.h"
#include "spl_rules.cuh"

struct spl_ctx {
    int n;
    SplRules rules;
    int device;
    int use_tma;
    int sm_count;
};

int spl_fail_(int code, const char* what, cudaError_t e = cudaSuccess);   // records spl_last_error() text, returns code
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return spl_fail_(SPL_E_CUDA, #call, e_); } while (0)

// spl_nnet.cu: the fused evaluator on rows that live in one of two places (row_src[r] = 0: states / valids bytes, 1: alt_states
// with alt_stride bytes per row + 13 mask words at alt_mask[w * alt_mask_stride + r]); row_src == NULL: all rows from states / valids.
// programmatic_dependent: launched as a programmatic dependent of the previous kernel in `st` (its set-up overlaps that kernel's tail)
int spl_nnet_forward_rows_(spl_ctx* c, const void* blob, const int8_t* states, const uint8_t* valids, const uint8_t* row_src,
                           const int8_t* alt_states, int alt_stride, const uint32_t* alt_mask, int alt_mask_stride, int n_rows, float* pi,
                           float* v, cudaStream_t st, bool programmatic_dependent = false);

// spl_nnet2.cu: the transposed evaluator (default); spl_nnet.cu keeps the first version behind SPL_NNET_IMPL=1 for comparison
namespace nn2 {
size_t blob_bytes(int n);
int pack(int n, const float* const* T, void* blob, size_t blob_bytes_);
int forward_rows(spl_ctx* c, const void* blob, const int8_t* states, const uint8_t* valids, const uint8_t* row_src, const int8_t* alt_states,
                 int alt_stride, const uint32_t* alt_mask, int alt_mask_stride, int n_rows, float* pi, float* v, cudaStream_t st,
                 bool programmatic_dependent);
int debug_stamps(long long* out32);
int debug_tile_stamps(long long* out);
}
int spl_nnet_impl_();   // 1 or 2 (environment variable SPL_NNET_IMPL, read once)

#define DISPATCH_N(n, ...)                                   \
    switch (n) {                                             \
        case 2: { constexpr int N = 2; __VA_ARGS__; } break; \
        case 3: { constexpr int N = 3; __VA_ARGS__; } break; \
        default: { constexpr int N = 4; __VA_ARGS__; } break;\
    }
