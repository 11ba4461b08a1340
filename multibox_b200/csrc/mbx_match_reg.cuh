// multibox_b200 -- register-resident matching + loss kernel (sm_100a).
//
// Same algorithm, arithmetic and outputs as the generic kernel in mbx_match.cu
// (see the header comment there for the reference mapping: loss.py:8-53,
// 55-117, model.py:322), restructured for latency and instruction count:
//
//   * each thread OWNS C columns (priors) j = tid + c*T and keeps their whole
//     solver state in registers: absolute box (4), log terms (2), confidence (1),
//     dual v (fp64), shortest-path cost (fp64), assigned row, path tag.  The
//     Dijkstra scan is a fully unrolled loop over C independent columns (ILP),
//     with no shared-memory traffic besides the broadcast of the scanned GT row;
//   * the first Dijkstra step of every augmentation (row `cur`: min_val = 0,
//     u[cur] = 0, every column unscanned) is specialised: r = C(cur, j) - v[j];
//   * the block-wide arg-min works on order-preserving 64-bit integer images of
//     the fp64 path costs with redux.sync (3 warp reductions + 1 vote per
//     stage, two stages, ONE __syncthreads per Dijkstra step); an exact tie at
//     the minimum diverts to the slow path that applies scipy's scan-order rule;
//   * the augmenting path is recovered from a tiny per-augmentation log
//     (removed column, its position in scipy's `remaining` list, the visit
//     index of the row that reached it) instead of a per-column path array;
//   * priors are staged once per CTA by a TMA bulk copy (cp.async.bulk).
//
// Shared memory per CTA: priors 16P + row4col 2P + O(M) -> ~13 KB at P=646.
#pragma once
#include <cooperative_groups.h>

#include "mbx_match.cuh"

// Optional phase timing (profiles/phase_timing.py builds with -DMBX_PHASE_TIMING): per-warp cycle
// totals of the Dijkstra-step phases, written to the mask output buffer.  Compiled out otherwise.
#ifdef MBX_PHASE_TIMING
#define MBX_T(k)                                         \
    do {                                                 \
        const long long t_now__ = clock64();             \
        t_acc[k] += t_now__ - t_last;                    \
        t_last = t_now__;                                \
    } while (0)
#else
#define MBX_T(k) \
    do {         \
    } while (0)
#endif

namespace mbx {

namespace {

constexpr unsigned kPayNone = 0xffffffffu;

struct RSmem {
    float4 *priors, *gt;
    double *u, *red;
    int4 *part;                 // [2][NWARPS] {key_hi, key_lo, payload, -}
    unsigned long long *pk;     // [NWARPS]
    int *col4row, *rm_col, *rm_idx, *rm_pm, *visit, *ri, *ctl;   // ctl[0] next general row, ctl[1] fast path off
    uint2 *rowpart;             // [M][NWARPS] per-warp first-step minimum of each row {key32, column | tie<<31}
    uint2 *rowmin;              // [M] block-wide first-step minimum of each row
    short *row4col;
    // per-column state of the general search (touched only for conflict / tie rows, so it lives
    // in shared memory and the hot first-step pass keeps the registers)
    double *cv;                 // [P] column dual v
    double *spc;                // [P] shortest path cost when it is not simply (double)c0 (bit in `dbl`)
    float *c0;                  // [P] cost of the column against the row being searched (first step)
    short *pmv;                 // [P] visit index of the row that set spc (bit in `updm`)
    short *arow;                // [P] row the column was assigned to when it was scanned
    unsigned char *dirty;       // [P] column dual is non-zero
    uint64_t *bar;
};

__host__ __device__ inline size_t rcarve(RSmem *s, unsigned char *base, int P, int M, int nwarps, bool has_priors) {
    size_t o = 0;
    auto take = [&](size_t bytes, size_t al) {
        o = align_up(o, al);
        size_t r = o;
        o += bytes;
        return r;
    };
    const int Mx = M > 0 ? M : 1;
    size_t o_pri = take(has_priors ? sizeof(float4) * P : 0, 16);
    size_t o_gt = take(sizeof(float4) * Mx, 16);
    size_t o_part = take(sizeof(int4) * 2 * nwarps, 16);
    size_t o_rp = take(sizeof(uint2) * static_cast<size_t>(Mx) * nwarps, 8);
    size_t o_rm = take(sizeof(uint2) * Mx, 8);
    size_t o_cv = take(sizeof(double) * P, 8);
    size_t o_spc = take(sizeof(double) * P, 8);
    size_t o_u = take(sizeof(double) * Mx, 8);
    size_t o_red = take(sizeof(double) * 3 * nwarps, 8);
    size_t o_pk = take(sizeof(unsigned long long) * nwarps, 8);
    size_t o_bar = take(8, 8);
    size_t o_c4r = take(sizeof(int) * Mx, 4);
    size_t o_rmc = take(sizeof(int) * (M + 2), 4);
    size_t o_rmi = take(sizeof(int) * (M + 2), 4);
    size_t o_rmp = take(sizeof(int) * (M + 2), 4);
    size_t o_vis = take(sizeof(int) * (M + 2), 4);
    size_t o_ri = take(sizeof(int) * nwarps, 4);
    size_t o_ctl = take(sizeof(int) * 4, 4);
    size_t o_c0 = take(sizeof(float) * P, 4);
    size_t o_r4c = take(sizeof(short) * P, 2);
    size_t o_pmv = take(sizeof(short) * P, 2);
    size_t o_arow = take(sizeof(short) * P, 2);
    size_t o_dirty = take(P, 1);
    if (s) {
        s->priors = reinterpret_cast<float4 *>(base + o_pri);
        s->gt = reinterpret_cast<float4 *>(base + o_gt);
        s->part = reinterpret_cast<int4 *>(base + o_part);
        s->u = reinterpret_cast<double *>(base + o_u);
        s->red = reinterpret_cast<double *>(base + o_red);
        s->pk = reinterpret_cast<unsigned long long *>(base + o_pk);
        s->bar = reinterpret_cast<uint64_t *>(base + o_bar);
        s->col4row = reinterpret_cast<int *>(base + o_c4r);
        s->rm_col = reinterpret_cast<int *>(base + o_rmc);
        s->rm_idx = reinterpret_cast<int *>(base + o_rmi);
        s->rm_pm = reinterpret_cast<int *>(base + o_rmp);
        s->visit = reinterpret_cast<int *>(base + o_vis);
        s->ri = reinterpret_cast<int *>(base + o_ri);
        s->row4col = reinterpret_cast<short *>(base + o_r4c);
        s->rowpart = reinterpret_cast<uint2 *>(base + o_rp);
        s->rowmin = reinterpret_cast<uint2 *>(base + o_rm);
        s->ctl = reinterpret_cast<int *>(base + o_ctl);
        s->dirty = base + o_dirty;
        s->cv = reinterpret_cast<double *>(base + o_cv);
        s->spc = reinterpret_cast<double *>(base + o_spc);
        s->c0 = reinterpret_cast<float *>(base + o_c0);
        s->pmv = reinterpret_cast<short *>(base + o_pmv);
        s->arow = reinterpret_cast<short *>(base + o_arow);
    }
    return align_up(o, 16);
}

// order-preserving map double -> uint64 (-0.0 and +0.0 share one image)
__device__ __forceinline__ unsigned long long ord64(double x) {
    long long b = __double_as_longlong(x);
    unsigned long long k = static_cast<unsigned long long>(b) ^
                           (static_cast<unsigned long long>(b >> 63) | 0x8000000000000000ull);
    return k + (k == 0x7fffffffffffffffull);
}
__device__ __forceinline__ double unord64(unsigned long long k) {
    const unsigned long long b = (k & 0x8000000000000000ull) ? (k ^ 0x8000000000000000ull) : ~k;
    return __longlong_as_double(static_cast<long long>(b));
}

// order-preserving map float -> uint32 (-0.0 and +0.0 share one image)
__device__ __forceinline__ unsigned ord32(float x) {
    const unsigned b = __float_as_uint(__fadd_rn(x, 0.0f));
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unord32(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
constexpr unsigned kColNone = 0x7fffffffu;
constexpr unsigned kOrdInf32 = 0xff800000u;   // ord32(+inf)

// lexicographic min of (key, col) over a warp; `tie` <=> two different columns share the minimal key
__device__ __forceinline__ void warp_rowmin(unsigned &key, unsigned &col, bool &tie) {
    const unsigned mk = __reduce_min_sync(0xffffffffu, key);
    const bool mine = key == mk;
    const unsigned mc = __reduce_min_sync(0xffffffffu, mine ? col : kColNone);
    tie = __any_sync(0xffffffffu, mine && (tie || col != mc));
    key = mk;
    col = mc;
}

template <int NWARPS>
__device__ __forceinline__ void block_sync() {
    if (NWARPS == 1)
        __syncwarp();
    else
        __syncthreads();
}

// One reduction stage over a warp: lexicographic min of (hi, lo, pay); `tie` becomes true
// when two different entries share the minimal (hi, lo) or the winner carried a tie already.
__device__ __forceinline__ void warp_argmin(unsigned &hi, unsigned &lo, unsigned &pay, bool &tie) {
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned lo2 = (hi == mh) ? lo : 0xffffffffu;
    const unsigned ml = __reduce_min_sync(0xffffffffu, lo2);
    const bool mine = (hi == mh) && (lo == ml);
    const unsigned mp = __reduce_min_sync(0xffffffffu, mine ? pay : kPayNone);
    tie = __any_sync(0xffffffffu, mine && (tie || pay != mp));
    hi = mh;
    lo = ml;
    pay = mp;
}

}  // namespace

// Per-column state of the GENERAL search (dual v, path cost, first-step cost, path tag, row at
// removal).  With few columns per thread (C <= 3: the wide, latency-oriented CTAs) it stays in
// registers; with many it lives in shared memory so that the hot first-step pass keeps the
// registers (more CTAs per SM).  Each column is only ever touched by its owner thread.
template <typename TV, int C, bool IN_REGS>
struct ColState {
    TV r[IN_REGS ? C : 1];
    __device__ __forceinline__ TV get(int c, int j, const TV *m) const {
        if constexpr (IN_REGS) {
            TV v = r[0];
#pragma unroll
            for (int q = 1; q < C; ++q) v = (q == c) ? r[q] : v;
            return v;
        } else {
            return m[j];
        }
    }
    __device__ __forceinline__ void set(int c, int j, TV *m, TV v) {
        if constexpr (IN_REGS) {
#pragma unroll
            for (int q = 0; q < C; ++q)
                if (q == c) r[q] = v;
        } else {
            m[j] = v;
        }
    }
};

// Register budget: aim at >= 16 resident warps per SM (<= 128 registers per thread) while a
// thread owns few columns; wide per-thread footprints (C >= 5) trade occupancy for registers.
template <int NWARPS, int C>
constexpr int min_blocks_per_sm() {
    const int warps_per_sm = (C == 4) ? 24 : ((C <= 3) ? 16 : ((C <= 6) ? 12 : 8));
    return (NWARPS >= warps_per_sm) ? 1 : warps_per_sm / NWARPS;
}

// CL > 1: a thread-block CLUSTER of CL CTAs solves one image (few-image, latency-bound
// batches: the columns, i.e. the n*P cost evaluations, are spread over CL SMs).  Column j is
// owned by cluster thread gtid = rank*T + tid with j = gtid + c*T*CL.  The small per-image
// tables every CTA reads (GT, u, row4col, col4row, the removal log, control words) are
// REPLICATED in each CTA's shared memory; whoever updates them stores to all replicas through
// distributed shared memory, and cluster barriers replace the CTA barriers where such updates
// must be visible.  Per-row first-step minima and the dirty marks live in the leader (rank 0).
//
// RSP > 1 (row split; CL == 1): the CTA holds RSP warp GROUPS that each own ALL the columns (thread
// `ctid` of every group holds the same C columns, loads and logs computed redundantly).  The groups
// share the batched first step by ROWS -- the longest phase of a latency-bound image is cut by RSP
// -- after which only group 0 carries column state; the helper groups take part in the barriers and
// block-wide reductions with empty candidates.  Chosen when every image has an SM to itself.
template <int NWARPS, int C, int CL, int RSP>
__global__ void __launch_bounds__(NWARPS * 32, (CL > 1 || RSP > 1) ? 1 : min_blocks_per_sm<NWARPS, C>())
mbx_match_loss_reg_kernel(const MatchParams p) {
    namespace cg = cooperative_groups;
    constexpr int T = NWARPS * 32;
    constexpr int TG = T / RSP;       // threads of one column group
    constexpr int TC = TG * CL;       // column stride of a thread
    constexpr int NPART = NWARPS * CL;   // warps per image
    constexpr int NPG = NPART / RSP;     // warps that share one row of the batched first step
    static_assert(NPART <= 32, "one lane per warp partial");
    static_assert(RSP == 1 || (CL == 1 && NWARPS % RSP == 0), "row split: single CTA, whole warp groups");
    constexpr bool RS = (C <= 3);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RSmem s;
    const bool boundary = (p.flags & MBX_FLAG_BOUNDARY) != 0;
    const bool logits = (p.flags & MBX_FLAG_LOGITS) != 0;
    const bool has_priors = !boundary;
    rcarve(&s, smem_raw, p.P, p.M, NPART, has_priors);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int crank = 0;
    if constexpr (CL > 1) crank = static_cast<int>(cg::this_cluster().block_rank());
    const int rgrp = (RSP > 1) ? tid / TG : 0;           // warp group (row split)
    const bool helper = RSP > 1 && rgrp != 0;            // holds columns for the first step only
    const int gtid = crank * TG + (RSP > 1 ? tid % TG : tid);   // column-owner index within the image's cluster
    const int gwarp = crank * NWARPS + warp;
    // barrier over every thread working on the image
    auto image_sync = [&]() {
        if constexpr (CL > 1)
            cg::this_cluster().sync();
        else
            block_sync<NWARPS>();
    };
    // store to every CTA's replica / to the leader's copy of a shared-memory location
    auto store_all = [&](auto *ptr, auto val) {
        if constexpr (CL > 1) {
#pragma unroll
            for (int r = 0; r < CL; ++r) *cg::this_cluster().map_shared_rank(ptr, r) = val;
        } else {
            *ptr = val;
        }
    };
    auto store_leader = [&](auto *ptr, auto val) {
        if constexpr (CL > 1)
            *cg::this_cluster().map_shared_rank(ptr, 0) = val;
        else
            *ptr = val;
    };
    const int P = p.P, M = p.M;
    const float half_alpha = __fdiv_rn(p.alpha, 2.0f);   // (alpha / 2.) in fp32, loss.py:35
    const double INF = CUDART_INF;
    unsigned status = 0;

    if constexpr (CL > 1) cg::this_cluster().sync();   // every CTA of the cluster runs before any remote shared-memory access
    __shared__ HeadTab sh_heads[MBX_MAX_HEADS];
    const int nheads = p.nheads;
    if (nheads > 1) stage_heads(p, sh_heads);
    if (has_priors) {
        if (tid == 0) {
            mbar_init(s.bar, 1);
            fence_mbar_init();
        }
        block_sync<NWARPS>();
        if (tid == 0) {
            mbar_arrive_expect_tx(s.bar, static_cast<uint32_t>(sizeof(float4) * P));
            bulk_copy_g2s(s.priors, p.priors, static_cast<uint32_t>(sizeof(float4) * P), s.bar);
        }
    } else if (nheads > 1) {
        block_sync<NWARPS>();
    }
    bool priors_ready = !has_priors;
    int pbuf = 0;
#ifdef MBX_PHASE_TIMING
    long long t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t_last = clock64();
    unsigned long long t_g0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_g0));
#endif

    unsigned invalid_mask = 0;   // columns of this thread beyond P (helper groups: all, for the general search)
#pragma unroll
    for (int c = 0; c < C; ++c)
        if (helper || gtid + c * TC >= P) invalid_mask |= 1u << c;

    // Deferred fused all-reduce: the launch appends one extra CTA that only sends the previous
    // step's loss sums to the peers (NVLink latency overlaps this kernel's work).
    const bool has_poster = (CL == 1) && p.ar_world > 1 && (p.flags & MBX_FLAG_AR_DEFERRED);
    const int n_work = has_poster ? static_cast<int>(gridDim.x) - 1 : static_cast<int>(gridDim.x) / CL;
    const bool is_poster = has_poster && static_cast<int>(blockIdx.x) == n_work;
    if (is_poster && warp == 0) ar_post_pending(p);

    // Image scheduling.  Static (image = CTA index, stride = resident CTAs) when every image has
    // its own CTA; otherwise DYNAMIC over the heavy-first order built by mbx_order_kernel: the
    // first wave takes positions 0..n_work-1, later positions are claimed from a global counter.
    // Thread 0 issues the claim when an image starts and reads it when the image is done, so the
    // L2 round trip of the atomic is off the critical path.  (Making the order kernel a programmatic
    // dependency -- first wave in index order, griddepcontrol.wait before order[] is first read --
    // was measured: it hides the 4-5 us order kernel but gives up heavy-first for the first wave,
    // which costs as much on skewed batches; not kept.)
    const bool dyn = (CL == 1) && p.dynamic != 0;
    for (int q = is_poster ? p.B : static_cast<int>(blockIdx.x) / CL; q < p.B;) {
        const int b = dyn ? __ldcg(p.order + q) : q;
        unsigned claim = 0u;
        if (dyn && tid == 0) claim = atomicAdd(p.queue, 1u);
        const float4 *gg;
        int n = image_gt(p, b, gg);
        if (n < 0 || n > M) {
            status |= MBX_STATUS_BAD_NUM_GT;
            n = n < 0 ? 0 : M;
        }
        const size_t row0 = static_cast<size_t>(b) * P;
        if (!priors_ready) {
            mbar_wait(s.bar, 0);
            priors_ready = true;
        }
        // ---- per-column state in registers
        float4 loc[C];
        float lc[C], l1[C], cf[C];
        ColState<double, C, RS> cv, spc;    // dual v; path cost when it is not simply (double)c0 (bit in `dbl`)
        ColState<float, C, RS> c0;          // first-step cost against the row being searched
        ColState<short, C, RS> ptag, arow;   // path tag (bit in `updm`); row the column had when scanned
        unsigned vnz = 0u;    // which of this thread's columns have a non-zero dual
        const float4 *gl = reinterpret_cast<const float4 *>(p.locations) + row0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = gtid + c * TC;
            loc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            cf[c] = 0.5f;
            if (j < P) {
                if (nheads > 1) {   // straight from the per-head conv outputs (model.py:295-320 never materialised)
                    int hh;
                    const size_t e = head_elem(sh_heads, nheads, j, b, hh);
                    loc[c] = ld_stream_f4(reinterpret_cast<const float4 *>(sh_heads[hh].loc) + e);
                    cf[c] = ld_stream_f(sh_heads[hh].conf + e);
                } else {
                    loc[c] = ld_stream_f4(gl + j);
                    cf[c] = ld_stream_f(p.confidences + row0 + j);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = gtid + c * TC;
            if (RS || j < P) {
                cv.set(c, j, s.cv, 0.0);
                arow.set(c, j, s.arow, static_cast<short>(-1));
            }
            if (j < P) {
                if (has_priors) {
                    const float4 q = s.priors[j];
                    loc[c].x = __fadd_rn(loc[c].x, q.x);   // loss.py:71
                    loc[c].y = __fadd_rn(loc[c].y, q.y);
                    loc[c].z = __fadd_rn(loc[c].z, q.z);
                    loc[c].w = __fadd_rn(loc[c].w, q.w);
                }
                if (logits) {
                    cf[c] = sigmoidf_(cf[c]);              // model.py:322
                    if (p.conf_out && !helper) p.conf_out[row0 + j] = cf[c];
                }
                const float ce = boundary ? cf[c] : __fadd_rn(cf[c], kEps32);   // loss.py:74
                lc[c] = n > 0 ? nplogf(ce) : 0.0f;   // loss.py:21 (only ever read for an image that has GT rows)
                float w = __fsub_rn(1.0f, ce);                                   // loss.py:22-24
                if (w > 1.0f) w = 1.0f;
                if (w <= 0.0f) w = kEps32;
                l1[c] = nplogf(w);                                               // loss.py:25
            } else {
                lc[c] = -CUDART_INF_F;   // a column that does not exist costs +inf: never selected
                l1[c] = 0.0f;
            }
        }
        for (int i = tid; i < n; i += T) {
            s.gt[i] = gg[i];
            s.u[i] = 0.0;
            s.col4row[i] = -1;
        }
        for (int j = tid; j < P; j += T) {   // this CTA's replica of the per-column tables
            s.row4col[j] = -1;
            s.dirty[j] = 0;
        }
        if (tid == 0) s.ctl[1] = 0;
        block_sync<NWARPS>();
        MBX_T(0);   // prologue (load, logs)

        // ---- one shortest augmenting path per GT row (rows = GT, columns = priors)
        bool failed = false;
        bool ok = true;   // every cost entry seen so far is neither NaN nor -inf

        // ---- batched first Dijkstra step of EVERY row, assuming all column duals are zero.
        // Row i's first step is argmin_j (C(i,j) - v[j]); v is zero until an augmenting path
        // passes THROUGH a column, so all rows can be evaluated up front with no barrier and
        // RB*C independent cost chains per thread.  fp32 keys are exact here (r == C).
        {
            constexpr int RB = (C <= 2) ? 4 : ((C <= 3) ? 3 : 2);
            for (int i0 = rgrp * RB; i0 < n; i0 += RB * RSP) {
                float4 g[RB];
                float best[RB];
                unsigned bcol[RB], btie[RB];
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    g[r] = s.gt[(i0 + r < n) ? (i0 + r) : (n - 1)];
                    best[r] = CUDART_INF_F;
                    bcol[r] = kColNone;
                    btie[r] = 0u;
                }
#pragma unroll
                for (int c = 0; c < C; ++c) {
#pragma unroll
                    for (int r = 0; r < RB; ++r) {
                        const float c32 = cost32(loc[c], g[r], half_alpha, lc[c], l1[c]);
                        ok = ok && (c32 > -CUDART_INF_F);
                        const bool lt = c32 < best[r];
                        const unsigned eq = c32 == best[r] ? 1u : 0u;
                        best[r] = lt ? c32 : best[r];
                        bcol[r] = lt ? static_cast<unsigned>(gtid + c * TC) : bcol[r];
                        btie[r] = lt ? 0u : (btie[r] | eq);
                    }
                }
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    unsigned key = (bcol[r] == kColNone) ? 0xffffffffu : ord32(best[r]);
                    unsigned col = bcol[r];
                    bool tieb = btie[r] != 0u;
                    warp_rowmin(key, col, tieb);
                    if (lane == 0 && i0 + r < n) {
                        const uint2 e = make_uint2(key, col | (tieb ? 0x80000000u : 0u));
                        if (NPG > 1)
                            store_leader(&s.rowpart[(i0 + r) * NPG + gwarp % NPG], e);
                        else
                            s.rowmin[i0 + r] = e;
                    }
                }
            }
            if (NPG > 1) {
                image_sync();
                for (int i = warp; i < n && crank == 0; i += NWARPS) {
                    uint2 e = make_uint2(0xffffffffu, kColNone);
                    if (lane < NPG) e = s.rowpart[i * NPG + lane];
                    unsigned key = e.x, col = e.y & kColNone;
                    bool tieb = (e.y >> 31) != 0u;
                    warp_rowmin(key, col, tieb);
                    if (lane == 0) s.rowmin[i] = make_uint2(key, col | (tieb ? 0x80000000u : 0u));
                }
            }
            block_sync<NWARPS>();
        }
        MBX_T(1);   // batched first step

        // ---- rows in order.  Thread 0 disposes of every row whose precomputed first step is
        // decisive (unique minimum at a column whose dual is still zero and which is unassigned:
        // that column is the sink, the path is the single edge, no dual changes besides
        // u[row] = min).  A dual can only make its column MORE expensive (v <= 0, enforced below),
        // so a zero-dual minimum stays the true minimum.  Any other row is solved by the whole CTA
        // with the general shortest-augmenting-path search below.
        int cur = 0;
        for (;;) {
            if (warp == 0 && crank == 0) {
                // Warp 0 (of the leader) disposes of up to 32 consecutive rows per round: lane r takes row cur+r.
                // A row is decisive when its first-step minimum is unique, finite, at a column with
                // zero dual that is unassigned AND not wanted by an earlier row of the same round.
                // The round commits the rows before the first non-decisive one.
                const bool fast_off = s.ctl[1] != 0;
                while (cur < n && !fast_off) {
                    const int row = cur + lane;
                    bool good = false;
                    unsigned col = kColNone;
                    uint2 rm = make_uint2(0u, 0u);
                    if (row < n) {
                        rm = s.rowmin[row];
                        col = rm.y & kColNone;
                        good = !(rm.y >> 31) && rm.x < kOrdInf32 && col != kColNone;
                        if (good) good = !s.dirty[col] && s.row4col[col] < 0;
                    }
                    // an earlier lane of this round wants the same column -> this row conflicts
                    const unsigned same = __match_any_sync(0xffffffffu, col);
                    if (good && (same & ((1u << lane) - 1u))) good = false;
                    const unsigned bad = __ballot_sync(0xffffffffu, !good);   // rows >= n are "bad" too
                    const int nfast = bad ? (__ffs(bad) - 1) : 32;
                    if (lane < nfast) {
                        store_all(&s.row4col[col], static_cast<short>(row));
                        store_all(&s.col4row[row], static_cast<int>(col));
                        store_all(&s.u[row], static_cast<double>(unord32(rm.x)));
                    }
                    cur += nfast;
                    __syncwarp();
                    if (nfast < 32) break;
                }
                if (lane == 0) store_all(&s.ctl[0], cur);
            }
            image_sync();
            cur = s.ctl[0];
            MBX_T(2);   // sequential fast rows
            if (cur >= n) break;
            // assigned bits of this thread's columns (thread 0 assigned sinks on its own)
            unsigned asg = 0u;
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (!((invalid_mask >> c) & 1u) && s.row4col[gtid + c * TC] >= 0) asg |= 1u << c;
            int i = cur, R = 0;
            double min_val = 0.0, ui = 0.0;
            unsigned scmask = invalid_mask;   // columns already scanned (or non-existent)
            unsigned dbl = 0u, updm = 0u;     // see spc64 / pm above
            // The first Dijkstra step of this row is already known when its precomputed first-step
            // minimum is unique and sits at a column whose dual is still zero: duals only raise
            // reduced costs (v <= 0), so that column is the strict minimum of C(cur, j) - v[j] over
            // all j, exactly what scanning row `cur` would select (the same argument as for the
            // decisive rows above; the column is assigned, or the row would have been decisive).
            // The scan, the block arg-min and the barrier of step 0 are skipped; the per-column
            // first-step costs that step 0 would have left behind are formed at the start of step 1.
            bool skip0 = false;
            if constexpr (CL == 1) {
                const uint2 rm0 = s.rowmin[cur];
                const unsigned col0 = rm0.y & kColNone;
                if (s.ctl[1] == 0 && !(rm0.y >> 31) && rm0.x < kOrdInf32 && col0 != kColNone && !s.dirty[col0]) {
                    const int r4c0 = s.row4col[col0];
                    if (r4c0 >= 0) {
                        skip0 = true;
                        min_val = static_cast<double>(unord32(rm0.x));
                        const int cstar0 = static_cast<int>(col0) / TC;
                        if (!helper && static_cast<int>(col0) - cstar0 * TC == gtid) {   // the column's owner logs the removal
#pragma unroll
                            for (int c = 0; c < C; ++c)
                                if (c == cstar0) arow.set(c, static_cast<int>(col0), s.arow, static_cast<short>(r4c0));
                            scmask |= 1u << cstar0;
                            s.rm_col[0] = static_cast<int>(col0);
                            s.rm_idx[0] = replay_pos(static_cast<int>(col0), 0, P, s.rm_idx);
                            s.rm_pm[0] = 0;
                            s.visit[1] = r4c0;
                        }
                        R = 1;
                        i = r4c0;
                        ui = s.u[i];
                    }
                }
            }
            for (;;) {
                const float4 g = s.gt[i];
                unsigned long long key;
                unsigned bj = kPayNone, tie = 0u;
                if (helper) {
                    key = ~0ull;   // a helper group has no column in the search: empty candidate
                } else if (R == 0) {
                    // First Dijkstra step: min_val = 0, u[cur] = 0, nothing scanned, so
                    // r = (0 + C) - 0 - v = C - v.  Where v == 0 (every column that was never
                    // passed through by an augmenting path) r is the fp32 cost itself, and fp32
                    // order == fp64 order of the widened values: no fp64 work on this path.
                    float best32 = CUDART_INF_F;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const float c32 = cost32(loc[c], g, half_alpha, lc[c], l1[c]);
                        ok = ok && (c32 > -CUDART_INF_F);
                        if (!((invalid_mask >> c) & 1u)) c0.set(c, gtid + c * TC, s.c0, c32);
                        const bool lt = c32 < best32;
                        const unsigned eq = c32 == best32 ? 1u : 0u;
                        best32 = lt ? c32 : best32;
                        bj = lt ? ((static_cast<unsigned>(gtid + c * TC) << 1) | ((asg >> c) & 1u)) : bj;
                        tie = lt ? 0u : (tie | eq);
                    }
                    double best = static_cast<double>(best32);
                    if (vnz) {   // rare: redo the thread-local minimum in fp64 with the duals
                        best = INF;
                        bj = kPayNone;
                        tie = 0u;
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            const int jc = ((invalid_mask >> c) & 1u) ? 0 : (gtid + c * TC);
                            double sp = ((invalid_mask >> c) & 1u) ? INF : static_cast<double>(c0.get(c, jc, s.c0));
                            if ((vnz >> c) & 1u) {
                                sp = __dsub_rn(sp, cv.get(c, jc, s.cv));
                                spc.set(c, jc, s.spc, sp);
                                dbl |= 1u << c;
                            }
                            const bool lt = sp < best;
                            const unsigned eq = sp == best ? 1u : 0u;
                            best = lt ? sp : best;
                            bj = lt ? ((static_cast<unsigned>(gtid + c * TC) << 1) | ((asg >> c) & 1u)) : bj;
                            tie = lt ? 0u : (tie | eq);
                        }
                    }
                    key = (bj == kPayNone) ? ~0ull : ord64(best);
                } else {
                    if (skip0 && R == 1) {   // what the skipped step 0 would have stored: C(cur, j) (- v[j] where v != 0)
                        const float4 g0 = s.gt[cur];
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            if ((invalid_mask >> c) & 1u) continue;
                            const int jc = gtid + c * TC;
                            const float c32_0 = cost32(loc[c], g0, half_alpha, lc[c], l1[c]);
                            c0.set(c, jc, s.c0, c32_0);
                            if ((vnz >> c) & 1u) {
                                spc.set(c, jc, s.spc, __dsub_rn(static_cast<double>(c32_0), cv.get(c, jc, s.cv)));
                                dbl |= 1u << c;
                            }
                        }
                    }
                    double best = INF;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const float c32 = cost32(loc[c], g, half_alpha, lc[c], l1[c]);
                        const int jc = ((invalid_mask >> c) & 1u) ? 0 : (gtid + c * TC);   // in-bounds index
                        const double r =
                            __dsub_rn(__dsub_rn(__dadd_rn(min_val, static_cast<double>(c32)), ui), cv.get(c, jc, s.cv));
                        const bool live = !((scmask >> c) & 1u);
                        // (a column that is not live -- scanned or non-existent -- never reads column state:
                        // the in-bounds dummy index of a non-existent column belongs to another thread)
                        double old = INF;
                        if (live) old = ((dbl >> c) & 1u) ? spc.get(c, jc, s.spc) : static_cast<double>(c0.get(c, jc, s.c0));
                        const bool upd = live && (r < old);
                        if (upd) {
                            spc.set(c, jc, s.spc, r);
                            ptag.set(c, jc, s.pmv, static_cast<short>(R));
                            dbl |= 1u << c;
                            updm |= 1u << c;
                        }
                        const double sp = live ? (upd ? r : old) : INF;
                        const bool lt = sp < best;
                        const unsigned eq = sp == best ? 1u : 0u;
                        best = lt ? sp : best;
                        bj = lt ? ((static_cast<unsigned>(gtid + c * TC) << 1) | ((asg >> c) & 1u)) : bj;
                        tie = lt ? 0u : (tie | eq);
                    }
                    key = (bj == kPayNone) ? ~0ull : ord64(best);
                }
                MBX_T(3);   // general scan
                // ---- block-wide arg-min of (path cost, column); exact ties flagged
                unsigned hi = static_cast<unsigned>(key >> 32), lo = static_cast<unsigned>(key);
                unsigned pay = bj;
                bool tieb = tie != 0u;
                warp_argmin(hi, lo, pay, tieb);
                if (NPART > 1) {
                    if (lane == 0)
                        store_all(&s.part[pbuf * NPART + gwarp],
                                  make_int4(static_cast<int>(hi), static_cast<int>(lo), static_cast<int>(pay), tieb ? 1 : 0));
                    image_sync();
                    int4 e = make_int4(-1, -1, -1, 0);
                    if (lane < NPART) e = s.part[pbuf * NPART + lane];
                    hi = static_cast<unsigned>(e.x);
                    lo = static_cast<unsigned>(e.y);
                    pay = static_cast<unsigned>(e.z);
                    tieb = e.w != 0;
                    warp_argmin(hi, lo, pay, tieb);
                    pbuf ^= 1;
                }
                MBX_T(4);   // block arg-min (stage 2)
                const unsigned long long mkey = (static_cast<unsigned long long>(hi) << 32) | lo;
                if (pay == kPayNone || mkey >= 0xfff0000000000000ull) {   // min is +inf: infeasible (scipy raises)
                    status |= MBX_STATUS_INFEASIBLE;
                    failed = true;
                    break;
                }
                min_val = unord64(mkey);
                int jstar = static_cast<int>(pay >> 1);
                bool is_sink = !(pay & 1u);   // an unassigned column ends the search
                if (tieb) {
                    // scipy's rule among the columns AT the minimum: the LAST unassigned one in
                    // `remaining` order wins, else the FIRST assigned one (rare path).
                    unsigned long long k = ~0ull;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        if ((scmask >> c) & 1u) continue;   // scanned or non-existent column
                        const double sp = ((dbl >> c) & 1u) ? spc.get(c, gtid + c * TC, s.spc) : static_cast<double>(c0.get(c, gtid + c * TC, s.c0));
                        if (!(sp == min_val)) continue;
                        const int j = gtid + c * TC;
                        const int pos = replay_pos(j, R, P, s.rm_idx);
                        const bool assigned = (asg >> c) & 1u;
                        const unsigned k2 = assigned ? static_cast<unsigned>(P + pos) : static_cast<unsigned>(P - 1 - pos);
                        const unsigned long long kk = (static_cast<unsigned long long>(k2) << 32) |
                                                      (static_cast<unsigned>(j) << 1) | (assigned ? 1u : 0u);
                        k = kk < k ? kk : k;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const unsigned long long t = __shfl_xor_sync(0xffffffffu, k, o);
                        k = t < k ? t : k;
                    }
                    if (NPART > 1) {
                        if (lane == 0) store_all(&s.pk[gwarp], k);
                        image_sync();
                        k = s.pk[0];
#pragma unroll
                        for (int w = 1; w < NPART; ++w) k = s.pk[w] < k ? s.pk[w] : k;
                        image_sync();
                    }
                    jstar = static_cast<int>((k & 0xffffffffu) >> 1);
                    is_sink = !(k & 1u);
                }
                const int cstar = jstar / TC;
                const bool owner = !helper && (jstar - cstar * TC) == gtid;
                if (is_sink) {
                    if (owner) {
                        // ---- the sink's owner augments along the path back to row `cur`.  Every log
                        // entry it reads was written before an earlier barrier; the sink itself needs
                        // no log entry (nothing is scanned after it).
                        int pmv = 0;
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            if (c == cstar && ((updm >> c) & 1u)) pmv = ptag.get(c, jstar, s.pmv);
                        scmask |= 1u << cstar;
                        asg |= 1u << cstar;
                        store_all(&s.u[cur], min_val);          // u[cur] was 0: 0 + min_val
                        int col = jstar, m = pmv;
                        for (;;) {
                            const int row = (m == 0) ? cur : s.visit[m];
                            store_all(&s.row4col[col], static_cast<short>(row));
                            store_all(&s.col4row[row], col);
                            if (m == 0) break;
                            col = s.rm_col[m - 1];
                            m = s.rm_pm[m - 1];
                        }
                    }
                    ++R;
                    break;
                }
                // Row of the assigned column jstar.  No walk can be in flight here: the previous
                // augmentation's walk finished before this step's barrier.
                const int r4c_star = s.row4col[jstar];
                if (owner) {   // remove jstar from the scan set and log it
                    int pmv = 0;
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        if (c == cstar) {
                            pmv = ((updm >> c) & 1u) ? ptag.get(c, jstar, s.pmv) : 0;
                            arow.set(c, jstar, s.arow, static_cast<short>(r4c_star));
                        }
                    scmask |= 1u << cstar;
                    store_all(&s.rm_col[R], jstar);
                    store_all(&s.rm_idx[R], replay_pos(jstar, R, P, s.rm_idx));
                    store_all(&s.rm_pm[R], pmv);
                    store_all(&s.visit[R + 1], r4c_star);
                }
                ++R;
                i = r4c_star;
                ui = s.u[i];
                if (NPART == 1) __syncwarp();
            }
            if (failed) break;
            MBX_T(5);   // selection, log, walk
            // ---- dual update: v (owner registers), u of the visited rows (shared, distinct rows).
            // Only columns scanned BEFORE the sink move (the sink's own delta is 0).
            if (R > 1) {
                const unsigned scanned = scmask & ~invalid_mask;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int j = gtid + c * TC;
                    if (!((scanned >> c) & 1u)) continue;
                    const int ar = arow.get(c, j, s.arow);
                    if (ar >= 0) {
                        const double sp = ((dbl >> c) & 1u) ? spc.get(c, j, s.spc) : static_cast<double>(c0.get(c, j, s.c0));
                        const double delta = __dsub_rn(min_val, sp);
                        const double vn = __dsub_rn(cv.get(c, j, s.cv), delta);
                        cv.set(c, j, s.cv, vn);
                        store_all(&s.u[ar], __dadd_rn(s.u[ar], delta));
                        if (vn != 0.0) {
                            vnz |= 1u << c;
                            store_leader(&s.dirty[gtid + c * TC], static_cast<unsigned char>(1));
                            // the precomputed first steps rely on v <= 0; fp rounding could in
                            // principle break that by an ulp: then every later row goes general
                            if (vn > 0.0) store_leader(&s.ctl[1], 1);
                        }
                        arow.set(c, j, s.arow, static_cast<short>(-1));
                    }
                }
            }
            ++cur;
            image_sync();   // walk, duals and dirty marks visible to the leader's warp 0
        }
        MBX_T(6);   // dual update
        if (!ok) status |= MBX_STATUS_INVALID_COST;
        block_sync<NWARPS>();   // the last walk's row4col / col4row are visible below

        // ---- epilogue: mask, matched GT index, loss terms, gradients
        double acc_sq = 0.0, acc_conf = 0.0;
        int n_match = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = gtid + c * TC;
            if (j >= P || helper) continue;
            const int r = s.row4col[j];
#ifndef MBX_PHASE_TIMING   // (timing builds use the mask buffer for the cycle counters)
            if (p.mask) p.mask[row0 + j] = r >= 0 ? 1 : 0;
#endif
            if (p.gt_idx) p.gt_idx[row0 + j] = r;
            n_match += r >= 0;
            const float ce = boundary ? cf[c] : __fadd_rn(cf[c], kEps32);
            float4 dl = make_float4(0.f, 0.f, 0.f, 0.f);
            float dc;
            if (r >= 0) {
                const float4 g = s.gt[r];
                const float d0 = __fsub_rn(loc[c].x, g.x), d1 = __fsub_rn(loc[c].y, g.y),
                            d2 = __fsub_rn(loc[c].z, g.z), d3 = __fsub_rn(loc[c].w, g.w);
                acc_sq += static_cast<double>(__fmul_rn(d0, d0));
                acc_sq += static_cast<double>(__fmul_rn(d1, d1));
                acc_sq += static_cast<double>(__fmul_rn(d2, d2));
                acc_sq += static_cast<double>(__fmul_rn(d3, d3));
                dl = make_float4(__fmul_rn(p.alpha, d0), __fmul_rn(p.alpha, d1), __fmul_rn(p.alpha, d2),
                                 __fmul_rn(p.alpha, d3));
                acc_conf -= static_cast<double>(lc[c]);
                dc = __fdiv_rn(-1.0f, ce);
            } else {
                const float one_m = __fsub_rn(1.0f, ce);
                const float arg = __fadd_rn(one_m, kEps32);   // loss.py:101
                float vcl = one_m;
                if (vcl > 1.0f) vcl = 1.0f;
                if (vcl <= 0.0f) vcl = kEps32;
                const float la = (arg == vcl) ? l1[c] : nplogf(arg);
                acc_conf -= static_cast<double>(la);
                dc = __fdiv_rn(1.0f, arg);
            }
            if (logits) dc = __fmul_rn(dc, __fmul_rn(cf[c], __fsub_rn(1.0f, cf[c])));
            if (nheads > 1) {   // gradients in the per-head layouts too
                int hh;
                const size_t e = head_elem(sh_heads, nheads, j, b, hh);
                if (sh_heads[hh].dloc) st_stream_f4(reinterpret_cast<float4 *>(sh_heads[hh].dloc) + e, dl);
                if (sh_heads[hh].dconf) sh_heads[hh].dconf[e] = dc;
            } else {
                if (p.d_loc) st_stream_f4(reinterpret_cast<float4 *>(p.d_loc) + row0 + j, dl);
                if (p.d_conf) p.d_conf[row0 + j] = dc;
            }
        }
        if (p.stacked && !failed && crank == 0) {
            const int off = p.stk_offsets[b];
            for (int i = tid; i < n; i += T) {
                const int pi = s.col4row[i];
                int rank = 0;
                for (int q = 0; q < n; ++q) rank += s.col4row[q] < pi;
                reinterpret_cast<float4 *>(p.stacked)[off + rank] = s.gt[i];
            }
        }
        acc_sq = warp_sum(acc_sq);
        acc_conf = warp_sum(acc_conf);
        n_match = __reduce_add_sync(0xffffffffu, n_match);
        if (lane == 0) {
            s.red[warp] = acc_sq;
            s.red[NWARPS + warp] = acc_conf;
            s.ri[warp] = n_match;
        }
        block_sync<NWARPS>();
        if (tid == 0) {
            double a = 0.0, cc = 0.0;
            int m = 0;
            for (int w = 0; w < NWARPS; ++w) {
                a += s.red[w];
                cc += s.red[NWARPS + w];
                m += s.ri[w];
            }
            p.partials[2 * (b * CL + crank)] = a;
            p.partials[2 * (b * CL + crank) + 1] = cc;
            p.img_matched[b * CL + crank] = m;
        }
        if (dyn && tid == 0) s.ctl[2] = n_work + static_cast<int>(claim);
        block_sync<NWARPS>();   // shared state is reused by the next image
        q = dyn ? s.ctl[2] : q + n_work;
    }

    if (status) atomicOr(p.status, status);
#ifdef MBX_PHASE_TIMING
    MBX_T(7);   // epilogue
    if (lane == 0 && p.mask) {
        long long *dbg = reinterpret_cast<long long *>(p.mask) + (static_cast<size_t>(blockIdx.x) * NWARPS + warp) * 10;
        for (int k = 0; k < 8; ++k) dbg[k] = t_acc[k];
        unsigned long long t_g1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_g1));
        dbg[8] = static_cast<long long>(t_g0);
        dbg[9] = static_cast<long long>(t_g1);
    }
#endif

    // ---- last CTA to finish reduces the per-image partials in a fixed order
    __shared__ bool is_last;
    __threadfence();
    block_sync<NWARPS>();
    if (tid == 0) {
        const unsigned t = atomicAdd(p.ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    block_sync<NWARPS>();
    if (!is_last) return;
    __threadfence();
    // status word and previous launch sequence number: loaded together with the partials
    const TailPrefetch pre = tail_prefetch(p);
    double a = 0.0, cc = 0.0, md = 0.0;
    for (int b = tid; b < p.B * CL; b += T) {
        a += __ldcg(p.partials + 2 * b);
        cc += __ldcg(p.partials + 2 * b + 1);
        md += static_cast<double>(__ldcg(p.img_matched + b));
    }
    a = warp_sum(a);
    cc = warp_sum(cc);
    md = warp_sum(md);
    if (lane == 0) {
        s.red[warp] = a;
        s.red[NWARPS + warp] = cc;
        s.red[2 * NWARPS + warp] = md;
    }
    block_sync<NWARPS>();
    if (warp == 0) {
        if (!has_poster && p.ar_world > 1 && (p.flags & MBX_FLAG_AR_DEFERRED)) ar_post_pending(p);
        double A = 0.0, Cc = 0.0, Mt = 0.0;
        for (int w = 0; w < NWARPS; ++w) {
            A += s.red[w];
            Cc += s.red[NWARPS + w];
            Mt += s.red[2 * NWARPS + w];
        }
        finalize_losses(p, A, Cc, Mt, pre);   // warp-cooperative (lane r posts to peer r)
    }
}

namespace {

struct KernelInfo {
    size_t configured_smem = 0;
    int occ = 0;
    size_t occ_smem = 0;
};

template <int NWARPS, int C, int CL, int RSP = 1>
int launch_one(const MatchParams &p, cudaStream_t st) {
    static thread_local KernelInfo info;
    auto kern = mbx_match_loss_reg_kernel<NWARPS, C, CL, RSP>;
    const size_t smem = rcarve(nullptr, nullptr, p.P, p.M, NWARPS * CL, !(p.flags & MBX_FLAG_BOUNDARY));
    if (smem > static_cast<size_t>(max_smem_optin())) return MBX_E_TOO_LARGE;
    if (smem > info.configured_smem) {
        if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    static_cast<int>(smem)),
                               "cudaFuncSetAttribute(match_reg)"))
            return e;
        info.configured_smem = smem;
        info.occ = 0;
    }
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(NWARPS * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CL > 1 ? 1 : 0;
    if (info.occ == 0 || info.occ_smem != smem) {
        if (CL > 1) {
            cfg.gridDim = dim3(CL * sm_count());
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess || nclusters < 1) {
                cudaGetLastError();
                return MBX_E_TOO_LARGE;   // clusters not schedulable: caller falls back to CL = 1
            }
            info.occ = nclusters;          // resident clusters on the whole device
        } else {
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&info.occ, kern, NWARPS * 32, smem);
            if (info.occ < 1) info.occ = 1;
            info.occ *= sm_count();        // resident CTAs on the whole device
        }
        info.occ_smem = smem;
    }
    int units = info.occ;                  // clusters (CL > 1) or CTAs
    if (units > p.B) units = p.B;
    const bool poster = (CL == 1) && p.ar_world > 1 && (p.flags & MBX_FLAG_AR_DEFERRED);
    cfg.gridDim = dim3(units * CL + (poster ? 1 : 0));
    MatchParams pp = p;
    if (CL == 1 && p.B > units && !(p.flags & MBX_FLAG_STATIC)) {
        // more images than resident CTAs: heavy-first order + dynamic scheduling
        if (int e = launch_order(p.num_gt, p.gt_row, 0, p.B, p.M, p.order, st)) return e;
        pp.dynamic = 1;
    }
    return check_cuda(cudaLaunchKernelEx(&cfg, kern, pp), "launch mbx_match_loss_reg_kernel");
}

}  // namespace

template <int NWARPS>
int launch_cols(const MatchParams &p, int cols, int cl, cudaStream_t st) {
    if (cl < 0) {   // row split: two warp groups that each hold every column (P <= 3 * 16 * NWARPS)
        if constexpr (NWARPS == 8 || NWARPS == 16) {
            switch (cols) {
                case 1: return launch_one<NWARPS, 1, 1, 2>(p, st);
                case 2: return launch_one<NWARPS, 2, 1, 2>(p, st);
                case 3: return launch_one<NWARPS, 3, 1, 2>(p, st);
                default: return MBX_E_TOO_LARGE;
            }
        }
        return MBX_E_TOO_LARGE;
    }
    if (cl > 1) {
        if constexpr (NWARPS == 8 || NWARPS == 16) {
            if (cl == 2 || (cl == 4 && NWARPS == 8)) {
                // cluster variants exist for thin per-thread footprints only
                if (cl == 2) {
                    switch (cols) {
                        case 1: return launch_one<NWARPS, 1, 2>(p, st);
                        case 2: return launch_one<NWARPS, 2, 2>(p, st);
                        case 3: return launch_one<NWARPS, 3, 2>(p, st);
                        default: return MBX_E_TOO_LARGE;
                    }
                } else {
                    if constexpr (NWARPS == 8) {
                        switch (cols) {
                            case 1: return launch_one<8, 1, 4>(p, st);
                            case 2: return launch_one<8, 2, 4>(p, st);
                            default: return MBX_E_TOO_LARGE;
                        }
                    }
                }
            }
        }
        return MBX_E_TOO_LARGE;
    }
    switch (cols) {
        case 1: return launch_one<NWARPS, 1, 1>(p, st);
        case 2: return launch_one<NWARPS, 2, 1>(p, st);
        case 3: return launch_one<NWARPS, 3, 1>(p, st);
        case 4: return launch_one<NWARPS, 4, 1>(p, st);
        case 5: return launch_one<NWARPS, 5, 1>(p, st);
        case 6: return launch_one<NWARPS, 6, 1>(p, st);
        case 7:
        case 8: return launch_one<NWARPS, 8, 1>(p, st);
        default: return MBX_E_TOO_LARGE;
    }
}

}  // namespace mbx
