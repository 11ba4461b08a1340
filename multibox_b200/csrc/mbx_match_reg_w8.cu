// multibox_b200 -- instantiations of the register-resident matching kernel for 8-warp CTAs.
#include "mbx_match_reg.cuh"

namespace mbx {
template int launch_cols<8>(const MatchParams &, int, cudaStream_t);
}  // namespace mbx
