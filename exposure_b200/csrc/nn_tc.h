// Internal interface between nn.cu (C-ABI entry points, CUDA-core engine) and nn_tc.cu
// (tcgen05 engine).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>

namespace expo {

enum { kBackendAuto = 0, kBackendSimt = 1, kBackendTcgen05 = 2 };
int gemm_backend();          // current setting (exp_set_gemm_backend)
bool use_tcgen05();          // resolves AUTO

bool tc_conv_fwd_supported(int Cout);
cudaError_t tc_conv_fwd(const float* x, int Cx, const float* vec, int Cv, float shift, const float* W,
                        const float* bias, const float* mask_ref, const float* post_mul, float* y, float* y2, int B,
                        int IH, int IW, int Cout, int mode, cudaStream_t st);
int tc_fc_splits(int M, int K, int N);
cudaError_t tc_fc_fwd_partials(const float* x, int ldx, const float* W, float* part, int M, int K, int N, int splits,
                               cudaStream_t st);

}  // namespace expo
