// multibox_b200 -- dispatch of the register-resident matching kernel family.  The kernel
// template lives in mbx_match_reg.cuh; its instantiations are compiled in parallel, one
// translation unit per CTA width (mbx_match_reg_w*.cu).
#include "mbx_match.cuh"

namespace mbx {

template <int NWARPS>
int launch_cols(const MatchParams &p, int cols, int cl, cudaStream_t st);
extern template int launch_cols<1>(const MatchParams &, int, int, cudaStream_t);
extern template int launch_cols<2>(const MatchParams &, int, int, cudaStream_t);
extern template int launch_cols<4>(const MatchParams &, int, int, cudaStream_t);
extern template int launch_cols<8>(const MatchParams &, int, int, cudaStream_t);
extern template int launch_cols<16>(const MatchParams &, int, int, cudaStream_t);

namespace {

int dispatch(const MatchParams &p, int nwarps, int cols, int cl, cudaStream_t st) {
    switch (nwarps) {
        case 1: return launch_cols<1>(p, cols, cl, st);
        case 2: return launch_cols<2>(p, cols, cl, st);
        case 4: return launch_cols<4>(p, cols, cl, st);
        case 8: return launch_cols<8>(p, cols, cl, st);
        case 16: return launch_cols<16>(p, cols, cl, st);
        default: return MBX_E_TOO_LARGE;
    }
}

}  // namespace

int launch_match_reg(const MatchParams &p, int force_warps, int force_cols, int force_cluster, cudaStream_t st) {
    if (p.P > 65535 || p.M > 32766) return MBX_E_TOO_LARGE;
    int nwarps = force_warps;
    if (nwarps == 0) {
        // CTA width: enough threads that a thread owns <= 3-4 columns (measured best for both the
        // few-image, latency-bound case and the many-image, throughput-bound case on B200).
        nwarps = p.P <= 96 ? 1 : (p.P <= 192 ? 2 : (p.P <= 384 ? 4 : (p.P <= 1024 ? 8 : 16)));
    }
    // One image per thread-block CLUSTER (2 or 4 SMs) is implemented and parity-tested, but on
    // B200 it measured no faster than one CTA per image even at B = 32 (cluster barriers and
    // remote stores eat what the split first-step pass saves: profiles/README.md), so it is only
    // used when forced through MBX_FLAG_CLUSTER_SHIFT.
    int cl = force_cluster ? force_cluster : 1;
    // Row split (two warp groups per image, each holding every column; the batched first step is
    // shared by rows), for shapes where a 256- or 128-thread group covers the priors with <= 3
    // columns per thread.  Implemented and parity-tested (bit-identical), but on B200 it measured
    // SLOWER than one group even at B = 32 (22.0 vs 20.5 us per launch on configs[1]: the redundant
    // loads / logs of the helper group and the 16-warp barriers cost more than the halved first step
    // saves), so it is only used when forced with MBX_FLAG_ROWSPLIT.
    const bool want_split = (p.flags & MBX_FLAG_ROWSPLIT) != 0 && !(p.flags & MBX_FLAG_NO_ROWSPLIT);
    if (want_split && cl == 1) {
        const int gw = force_warps ? force_warps / 2 : (p.P <= 384 ? 4 : 8);      // warps of one group
        const int cols = force_cols ? force_cols : (p.P + gw * 32 - 1) / (gw * 32);
        if ((gw == 4 || gw == 8) && cols >= 1 && cols <= 3 && cols * gw * 32 >= p.P) {
            const int rc = dispatch(p, 2 * gw, cols, -1, st);
            if (rc != MBX_E_TOO_LARGE) return rc;
        }
        if (p.flags & MBX_FLAG_ROWSPLIT) return MBX_E_TOO_LARGE;
    }
    if (cl > 1) {
        const int tc = nwarps * 32 * cl;
        const int cols = force_cols ? force_cols : (p.P + tc - 1) / tc;
        if (cols * tc >= p.P) {
            const int rc = dispatch(p, nwarps, cols, cl, st);
            if (rc != MBX_E_TOO_LARGE) return rc;
        }
        if (force_cluster) return MBX_E_TOO_LARGE;
    }
    int cols = force_cols ? force_cols : (p.P + nwarps * 32 - 1) / (nwarps * 32);
    while (!force_warps && cols > 8 && nwarps < 16) {
        nwarps *= 2;
        cols = (p.P + nwarps * 32 - 1) / (nwarps * 32);
    }
    if (cols * nwarps * 32 < p.P) return MBX_E_TOO_LARGE;
    return dispatch(p, nwarps, cols, 1, st);
}

}  // namespace mbx
