// attn_bwd_sm100.cu -- placeholder until the tcgen05 backward lands (next commit).
#include "common.cuh"
#include "internal.h"

namespace nnop {
bool attn_sm100_bwd_available() { return false; }
size_t attn_sm100_bwd_workspace_bytes(int, int, int, int) { return 0; }
int attn_sm100_bwd(const AttnParams&) {
  return fail(NNOP_ERR_ARG, "tcgen05 backward not built");
}
}  // namespace nnop
