// tcgen05 + TMA prefill kernel (placeholder until the kernel lands).
#include "attn_common.cuh"
namespace vattn {
bool prefill_tc_supported(const vattn_fwd_params_t&, std::string* why) {
  if (why) *why = "prefill tensor-core kernel not built yet";
  return false;
}
size_t prefill_tc_workspace(const vattn_fwd_params_t&) { return 0; }
void launch_prefill_tc(const vattn_fwd_params_t&, void*, size_t, cudaStream_t) {
  throw UnsupportedError("[vattn] prefill tensor-core kernel not built yet");
}
int run_umma_selftest(char* buf, size_t len, cudaStream_t) {
  if (buf && len) buf[0] = 0;
  return 0;
}
}  // namespace vattn
