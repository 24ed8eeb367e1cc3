// pmg_fast.cu — tcgen05 / TMA fast path for the committed PMG configuration (placeholder: not yet wired).
#include "pmg_common.cuh"

namespace pmg {

size_t fast_section_bytes(const Plan&) { return 0; }
bool fast_supported(const Plan&, const Shape&) { return false; }
size_t fast_workspace_bytes(const Plan&, const Shape&, int) { return 0; }
int fast_forward(const Plan&, const Shape&, const PackedLayout&, const char*, const void*, int, int, float*, float*,
                 const pmg_taps*, char*, cudaStream_t, StageTimer*) { return PMG_ERR_UNSUPPORTED; }
int fast_launch_count(const Plan&, const Shape&, int) { return 0; }
int pack_fast(const Plan&, const PackedLayout&, const float* const*, char*, cudaStream_t) { return PMG_OK; }
int gram_batched_fast(const void*, int, int, int, int, float, float*, cudaStream_t) { return PMG_ERR_UNSUPPORTED; }

}  // namespace pmg
