// multibox_b200 -- instantiations of the register-resident matching kernel for 4-warp CTAs.
#include "mbx_match_reg.cuh"

namespace mbx {
template int launch_cols<4>(const MatchParams &, int, cudaStream_t);
}  // namespace mbx
