// vq_tc.cu — tcgen05 nearest-code search (placeholder until the tensor-core kernel lands).
#include "common.cuh"
namespace vqb {
size_t vq_search_tc_workspace_bytes(const vqb_vq_desc*) { return 0; }
int vq_search_tc(const vqb_vq_desc*, const float*, const float*, const float*, const float*, int64_t*, void*, size_t,
                 cudaStream_t) {
  return set_err(VQB_ERR_UNIMPLEMENTED, "tensor-core VQ search is not built into this library");
}
}  // namespace vqb
