// tcgen05 + TMA decode kernel (placeholder until the kernel lands).
#include "attn_common.cuh"
namespace vattn {
bool decode_tc_supported(const vattn_fwd_params_t&, std::string* why) {
  if (why) *why = "decode tensor-core kernel not built yet";
  return false;
}
size_t decode_tc_workspace(const vattn_fwd_params_t&) { return 0; }
void launch_decode_tc(const vattn_fwd_params_t&, void*, size_t, cudaStream_t) {
  throw UnsupportedError("[vattn] decode tensor-core kernel not built yet");
}
}  // namespace vattn
