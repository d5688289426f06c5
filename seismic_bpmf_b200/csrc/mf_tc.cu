// mf_tc.cu -- tcgen05 (5th-gen tensor core) matched filter.  PLACEHOLDER: the tensor path is not
// wired yet; every shape resolves to the FP32 CUDA-core kernel.
#include "common.cuh"

namespace bpmf {
bool mf_tc_supported(size_t, size_t, int, int) { return false; }
size_t mf_tc_workspace_bytes(const MfShape &, int) { return 0; }
int mf_tc_run(const MfShape &, int, const float *, const int *, const float *, const float *, MfMeta,
              Workspace &, float *, size_t, cudaStream_t) {
    set_error("tensor-core matched filter is not available in this build");
    return BPMF_E_ARG;
}
}  // namespace bpmf
