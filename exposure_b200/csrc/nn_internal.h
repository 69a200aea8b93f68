// Internal interface between nn.cu (C-ABI entry points, exact-fp32 CUDA-core engine) and nn_tma.cu
// (TMA-fed tcgen05 engine).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>

namespace expo {

enum { kBackendAuto = 0, kBackendSimt = 1 };
int gemm_backend();          // current setting (exp_set_gemm_backend)
bool use_tma();              // TMA-fed tcgen05 engine where the shape allows it (the default)

}  // namespace expo

// nn_tma.cu: TMA-fed tcgen05 engine (backend 4); *_supported() decide per call, callers fall back
// to the CUDA-core engine for the shapes it does not take (first layer: Cin not a multiple of 32)
namespace expo {
bool tma_conv_fwd_supported(const float* x, int Cx, int Cv, float shift, const float* W, const float* bias,
                            const float* mask_ref, const float* post_mul, const float* y, const float* y2, int Cout);
cudaError_t tma_conv_fwd(const float* x, int Cx, const float* W, const float* bias, const float* mask_ref,
                         const float* post_mul, float* y, float* y2, int B, int IH, int IW, int Cout, int mode,
                         cudaStream_t st);
bool tma_conv_dgrad_supported(const float* dy, const float* W, const float* a_in, const float* dx, int Cin, int Cout);
cudaError_t tma_conv_dgrad(const float* dy, const float* W, const float* a_in, float* dx, int B, int IH, int IW, int Cin,
                           int Cout, cudaStream_t st);
bool tma_conv_wgrad_supported(const float* x, int Cx, int Cv, float shift, const float* dy, int Cout);
int tma_wgrad_splits(int B, int OH, int OW, int Cin, int Cout);
cudaError_t tma_conv_wgrad_partials(const float* x, int Cx, const float* dy, float* part, int B, int IH, int IW, int Cout,
                                    int splits, cudaStream_t st);
}  // namespace expo
