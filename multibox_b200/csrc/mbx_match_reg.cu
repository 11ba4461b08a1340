// multibox_b200 -- dispatch of the register-resident matching kernel family.  The kernel
// template lives in mbx_match_reg.cuh; its instantiations are compiled in parallel, one
// translation unit per CTA width (mbx_match_reg_w*.cu).
#include <unordered_map>

#include "mbx_match.cuh"

namespace mbx {

unsigned next_launch_id(const void *workspace_key) {
    static thread_local std::unordered_map<const void *, unsigned> launches;
    unsigned &c = launches[workspace_key];
    ++c;
    if (c == 0u) c = 1u;
    return c;
}

template <int NWARPS>
int launch_cols(const MatchParams &p, int cols, cudaStream_t st);
extern template int launch_cols<1>(const MatchParams &, int, cudaStream_t);
extern template int launch_cols<2>(const MatchParams &, int, cudaStream_t);
extern template int launch_cols<4>(const MatchParams &, int, cudaStream_t);
extern template int launch_cols<8>(const MatchParams &, int, cudaStream_t);
extern template int launch_cols<16>(const MatchParams &, int, cudaStream_t);

namespace {

int dispatch(const MatchParams &p, int nwarps, int cols, cudaStream_t st) {
    switch (nwarps) {
        case 1: return launch_cols<1>(p, cols, st);
        case 2: return launch_cols<2>(p, cols, st);
        case 4: return launch_cols<4>(p, cols, st);
        case 8: return launch_cols<8>(p, cols, st);
        case 16: return launch_cols<16>(p, cols, st);
        default: return MBX_E_TOO_LARGE;
    }
}

}  // namespace

// One CTA per image.  (Two ways of putting more hardware on one image -- a thread-block cluster
// per image over distributed shared memory, and a second warp group sharing the batched first
// step by rows -- were built and measured slower than one CTA in round 1; they were removed when
// the cheap-bound pruning made the first step a small part of the image: profiles/README.md.)
int launch_match_reg(const MatchParams &p, int force_warps, int force_cols, cudaStream_t st) {
    if (p.P > 65535 || p.M > 32766) return MBX_E_TOO_LARGE;
    int nwarps = force_warps;
    if (nwarps == 0) {
        // CTA width: enough threads that a thread owns <= 3-4 columns (measured best for both the
        // few-image, latency-bound case and the many-image, throughput-bound case on B200).
        // (up to 6 columns per thread an 8-warp CTA stays within 128 registers: two images per SM)
        nwarps = p.P <= 96 ? 1 : (p.P <= 192 ? 2 : (p.P <= 384 ? 4 : (p.P <= 1536 ? 8 : 16)));
    }
    int cols = force_cols ? force_cols : (p.P + nwarps * 32 - 1) / (nwarps * 32);
    while (!force_warps && cols > 8 && nwarps < 16) {
        nwarps *= 2;
        cols = (p.P + nwarps * 32 - 1) / (nwarps * 32);
    }
    if (cols * nwarps * 32 < p.P) return MBX_E_TOO_LARGE;
    return dispatch(p, nwarps, cols, st);
}

}  // namespace mbx
