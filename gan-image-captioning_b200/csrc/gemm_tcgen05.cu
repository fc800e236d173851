// tcgen05 / TMEM / TMA GEMM (placeholder until the tensor-core kernel lands): reports "not handled"
// so the dispatcher uses the exact-fp32 kernel.
#include "gic_internal.cuh"
namespace gic {
int gemm_tc(int, bool, bool, int, int, int, float, const float*, int, const float*, int, float, float*, int,
            const float*, cudaStream_t, bool* handled) {
  *handled = false;
  return GIC_OK;
}
}  // namespace gic
