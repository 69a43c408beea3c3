// Host-side builders shared by the stand-alone tensor-core launches and the POD launch.
#pragma once
#include <cuda.h>

#include "attn_tc_work.cuh"

namespace vattn {

struct DecodeTcLaunch {
  tcwork::DecodeTcParams dp;
  CUtensorMap kmap, vmap, kmap_tail, vmap_tail;
  SplitWorkspace ws;
};
// fills kernel parameters + TMA maps for a seqlen_q == 1 problem; `ws` receives the split partials
void build_decode_tc(const vattn_fwd_params_t& p, void* ws, cudaStream_t stream, DecodeTcLaunch* out,
                     bool unused = false);
// true when the decode kernel itself appends k_new/v_new (one new token per sequence)
bool decode_tc_fuses_append(const vattn_fwd_params_t& p);

struct PrefillTcLaunch {
  tcwork::PrefillParams pp;
  CUtensorMap qmap, kmap, vmap, kmap_tail, vmap_tail;
};
void build_prefill_tc(const vattn_fwd_params_t& p, PrefillTcLaunch* out);

bool decode_tc_supported(const vattn_fwd_params_t& p, std::string* why);
bool prefill_tc_supported(const vattn_fwd_params_t& p, std::string* why);
size_t decode_tc_workspace(const vattn_fwd_params_t& p);
size_t decode_tc_workspace_grid(const vattn_fwd_params_t& p);
void launch_combine(const vattn_fwd_params_t& p, int splits, const SplitWorkspace& ws, cudaStream_t stream);

}  // namespace vattn
