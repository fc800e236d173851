// GEMM dispatch: tcgen05 tensor-core kernels when the mode and the operand layout allow it,
// the exact-fp32 CUDA-core kernel otherwise.  Both are this library's own kernels; nothing here
// falls back to a CPU or to a vendor library.
#include "gic_internal.cuh"

namespace gic {

int gemm_tc(int mode, bool transA, bool transB, int M, int N, int K, float alpha, const float* A, int lda,
            const float* B, int ldb, float beta, float* C, int ldc, const float* bias, cudaStream_t stream,
            bool* handled);

int gemm(int mode, bool transA, bool transB, int M, int N, int K, float alpha, const float* A, int lda,
         const float* B, int ldb, float beta, float* C, int ldc, const float* bias, cudaStream_t stream) {
  if (mode != GEMM_FP32) {
    bool handled = false;
    int rc = gemm_tc(mode, transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, stream, &handled);
    if (rc != GIC_OK) return rc;
    if (handled) return GIC_OK;
  }
  return gemm_f32(transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, stream);
}

}  // namespace gic
