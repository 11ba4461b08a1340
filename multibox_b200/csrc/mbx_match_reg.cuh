// multibox_b200 -- register-resident matching + loss kernel (sm_100a).
//
// Same algorithm, arithmetic and outputs as the generic kernel in mbx_match.cu
// (see the header comment there for the reference mapping: loss.py:8-53,
// 55-117, model.py:322), restructured for latency and instruction count:
//
//   * each thread OWNS C columns (priors) j = tid + c*T and keeps their whole
//     solver state in registers: absolute box (4), log terms (2), confidence (1),
//     the per-prior constant of the cheap cost form (1), dual v (fp64),
//     shortest-path cost (fp64), assigned row, path tag;
//   * EXACT COSTS ONLY WHERE THEY CAN MATTER.  The reference's cost entry
//     (loss.py:35) takes ~36 dependent fp32 instructions in numpy's operation
//     order (cost32()).  mbx_bound.h gives a 4-FMA form a(i,j) + G_i with a
//     proven margin |c - (a + G)| <= m_j + mg_i.  Every place that needs a
//     minimum over columns first evaluates the cheap form, derives a threshold
//     from an upper bound of the answer, and evaluates cost32() only for the
//     columns whose lower bound does not exceed it -- typically one column per
//     row.  Whatever is skipped provably cannot be the minimum, tie for it, or
//     ever be selected, so every output bit is the one the full evaluation
//     produces (parity tests unchanged);
//   * the first Dijkstra step of EVERY row is batched up front (no column dual
//     is non-zero before an augmenting path passes through it): pass 1 = cheap
//     per-warp minima of all rows, pass 2 = exact costs of the few candidates;
//   * the general search keeps an upper bound UB of the final path cost (the
//     cheapest unassigned column seen so far) and scans a column exactly only
//     if its lower bound can beat UB;
//   * the block-wide arg-min works on order-preserving 64-bit integer images of
//     the fp64 path costs with redux.sync (3 warp reductions + 1 vote per
//     stage, two stages, ONE __syncthreads per Dijkstra step); an exact tie at
//     the minimum diverts to the slow path that applies scipy's scan-order rule;
//   * the augmenting path is recovered from a tiny per-augmentation log
//     (removed column, its position in scipy's `remaining` list, the visit
//     index of the row that reached it) instead of a per-column path array;
//   * priors are staged once per CTA by a TMA bulk copy (cp.async.bulk).
//
// Shared memory per CTA: priors 16P + row4col 2P + dirty P + O(M * warps) -> ~13 KB at P=646.
#pragma once
#include "mbx_bound.h"
#include "mbx_match.cuh"

// Optional phase timing (profiles/phase_timing.py builds with -DMBX_PHASE_TIMING): per-warp cycle
// totals of the phases, written to the mask output buffer.  Compiled out otherwise.
#ifdef MBX_PHASE_TIMING
#define MBX_T(k)                                         \
    do {                                                 \
        const long long t_now__ = clock64();             \
        t_acc[k] += t_now__ - t_last;                    \
        t_last = t_now__;                                \
    } while (0)
#define MBX_COUNT(k, v) t_acc[k] += (v)
#else
#define MBX_T(k) \
    do {         \
    } while (0)
#define MBX_COUNT(k, v) \
    do {                \
    } while (0)
#endif

namespace mbx {

namespace {

constexpr unsigned kPayNone = 0x7fffffffu;   // (bit 31 carries the tie flag in the cross-warp slot)
constexpr int kTimingSlots = 12;              // 10 accumulators + 2 global timestamps per warp

struct RSmem {
    float4 *priors, *gt, *gp;   // gp[i] = -2h * gt[i] (cheap cost form)
    float2 *rc;                 // [M] {G_i, mg_i}: row constant and row margin of the cheap form
    double *u, *red;
    int4 *part;                 // [2][NWARPS] {key_hi, key_lo, payload | tie << 31, bits of the unassigned minimum}
    unsigned long long *pk;     // [NWARPS]
    // first step, exact results per row: the FIRST candidate column to arrive (rcnt 0 -> 1) stores
    // (ord32(cost) << 32 | column) into rfirst with a plain store; later ones (rare: the candidate window
    // of a row usually holds one column) fold into rmin / rmax with 64-bit atomics
    unsigned long long *rfirst; // [M]
    unsigned long long *rmin;   // [M] min over the later candidates of (ord32(cost) << 32 | column)
    unsigned long long *rmax;   // [M] max over the later candidates of (~ord32(cost) << 32 | column)
    unsigned *rcnt;             // [M] candidates evaluated
    float *rowpart;             // [NWARPS][Mp] per-warp cheap first-step minimum of each row
    float *mw;                  // [NWARPS] per-warp margin of the cheap form
    int *col4row, *rm_col, *rm_idx, *rm_pm, *visit, *ri, *ctl;   // ctl[0] next general row, ctl[1] fast path off
    short *row4col;
    // per-column state of the general search when it does not fit the registers (C >= 4); each
    // column is only ever touched by its owner thread
    double *cv;                 // [P] column dual v
    double *spc;                // [P] shortest path cost
    short *pmv;                 // [P] visit index of the row that set spc
    short *arow;                // [P] row the column was assigned to when it was scanned
    unsigned char *dirty;       // [P] column dual is non-zero
    uint64_t *bar;
};

__host__ __device__ inline int rows_padded(int M) { return ((M > 0 ? M : 1) + 31) & ~31; }

__host__ __device__ inline size_t rcarve(RSmem *s, unsigned char *base, int P, int M, int nwarps, bool has_priors,
                                         bool col_state) {
    size_t o = 0;
    auto take = [&](size_t bytes, size_t al) {
        o = align_up(o, al);
        size_t r = o;
        o += bytes;
        return r;
    };
    const int Mx = M > 0 ? M : 1;
    const int Mp = rows_padded(M);
    const size_t Pc = col_state ? P : 0;
    size_t o_pri = take(has_priors ? sizeof(float4) * P : 0, 16);
    size_t o_gt = take(sizeof(float4) * Mx, 16);
    size_t o_gp = take(sizeof(float4) * Mx, 16);
    size_t o_part = take(sizeof(int4) * 2 * nwarps, 16);
    size_t o_rc = take(sizeof(float2) * Mx, 8);
    size_t o_rfirst = take(sizeof(unsigned long long) * Mx, 8);
    size_t o_rmin = take(sizeof(unsigned long long) * Mx, 8);
    size_t o_rmax = take(sizeof(unsigned long long) * Mx, 8);
    size_t o_cv = take(sizeof(double) * Pc, 8);
    size_t o_spc = take(sizeof(double) * Pc, 8);
    size_t o_u = take(sizeof(double) * Mx, 8);
    size_t o_red = take(sizeof(double) * 3 * nwarps, 8);
    size_t o_pk = take(sizeof(unsigned long long) * nwarps, 8);
    size_t o_bar = take(8, 8);
    size_t o_rp = take(sizeof(float) * static_cast<size_t>(Mp) * nwarps, 4);
    size_t o_mw = take(sizeof(float) * nwarps, 4);
    size_t o_rcnt = take(sizeof(unsigned) * Mx, 4);
    size_t o_c4r = take(sizeof(int) * Mx, 4);
    size_t o_rmc = take(sizeof(int) * (M + 2), 4);
    size_t o_rmi = take(sizeof(int) * (M + 2), 4);
    size_t o_rmp = take(sizeof(int) * (M + 2), 4);
    size_t o_vis = take(sizeof(int) * (M + 2), 4);
    size_t o_ri = take(sizeof(int) * nwarps, 4);
    size_t o_ctl = take(sizeof(int) * 4, 4);
    size_t o_r4c = take(sizeof(short) * P, 2);
    size_t o_pmv = take(sizeof(short) * Pc, 2);
    size_t o_arow = take(sizeof(short) * Pc, 2);
    size_t o_dirty = take(P, 1);
    if (s) {
        s->priors = reinterpret_cast<float4 *>(base + o_pri);
        s->gt = reinterpret_cast<float4 *>(base + o_gt);
        s->gp = reinterpret_cast<float4 *>(base + o_gp);
        s->part = reinterpret_cast<int4 *>(base + o_part);
        s->rc = reinterpret_cast<float2 *>(base + o_rc);
        s->rfirst = reinterpret_cast<unsigned long long *>(base + o_rfirst);
        s->rcnt = reinterpret_cast<unsigned *>(base + o_rcnt);
        s->rmin = reinterpret_cast<unsigned long long *>(base + o_rmin);
        s->rmax = reinterpret_cast<unsigned long long *>(base + o_rmax);
        s->u = reinterpret_cast<double *>(base + o_u);
        s->red = reinterpret_cast<double *>(base + o_red);
        s->pk = reinterpret_cast<unsigned long long *>(base + o_pk);
        s->bar = reinterpret_cast<uint64_t *>(base + o_bar);
        s->rowpart = reinterpret_cast<float *>(base + o_rp);
        s->mw = reinterpret_cast<float *>(base + o_mw);
        s->col4row = reinterpret_cast<int *>(base + o_c4r);
        s->rm_col = reinterpret_cast<int *>(base + o_rmc);
        s->rm_idx = reinterpret_cast<int *>(base + o_rmi);
        s->rm_pm = reinterpret_cast<int *>(base + o_rmp);
        s->visit = reinterpret_cast<int *>(base + o_vis);
        s->ri = reinterpret_cast<int *>(base + o_ri);
        s->row4col = reinterpret_cast<short *>(base + o_r4c);
        s->ctl = reinterpret_cast<int *>(base + o_ctl);
        s->dirty = base + o_dirty;
        s->cv = reinterpret_cast<double *>(base + o_cv);
        s->spc = reinterpret_cast<double *>(base + o_spc);
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

// minimum over the warp, NaNs dropped (SASS: CREDUX.MIN.F32, sm_100a)
__device__ __forceinline__ float warp_min_f32(float v) {
    float r;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
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

// First-step result of `row` after pass 2: key = ord32(minimum cost), col = lowest column at the
// minimum, unique = exactly one column attains it.  No candidate at all: key = 0xffffffff.
__device__ __forceinline__ void first_step_result(const RSmem &s, int row, unsigned &key, unsigned &col, bool &unique) {
    const unsigned cnt = s.rcnt[row];
    unsigned long long k1 = s.rfirst[row];
    if (cnt == 0u) k1 = ~0ull;
    unsigned long long k2 = ((k1 >> 32) ^ 0xffffffffull) << 32 | (k1 & 0xffffffffull);
    if (cnt > 1u) {
        const unsigned long long m1 = s.rmin[row], m2 = s.rmax[row];
        k1 = m1 < k1 ? m1 : k1;
        k2 = m2 > k2 ? m2 : k2;
    }
    key = static_cast<unsigned>(k1 >> 32);
    col = static_cast<unsigned>(k1);
    unique = cnt != 0u && static_cast<unsigned>(k2) == col;
}

__device__ __forceinline__ float bound_a(const float4 &l, const float4 &gq, float w) {
    return mbx_bound_a(l.x, l.y, l.z, l.w, gq.x, gq.y, gq.z, gq.w, w);
}

}  // namespace

// Per-column state of the GENERAL search (dual v, path cost, path tag, row at removal).  With few
// columns per thread (C <= 3: the wide, latency-oriented CTAs) it stays in registers; with many it
// lives in shared memory (more CTAs per SM).  Each column is only ever touched by its owner thread.
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
    const int warps_per_sm = (C == 4) ? 24 : ((C <= 6) ? 16 : 8);
    return (NWARPS >= warps_per_sm) ? 1 : warps_per_sm / NWARPS;
}

template <int NWARPS, int C>
__global__ void __launch_bounds__(NWARPS * 32, min_blocks_per_sm<NWARPS, C>())
mbx_match_loss_reg_kernel(const MatchParams p) {
    constexpr int T = NWARPS * 32;
    static_assert(NWARPS <= 32, "one lane per warp partial");
    static_assert(C <= 8, "candidate queue entries are row * 8 + c");
    constexpr bool RS = (C <= 3);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RSmem s;
    const bool boundary = (p.flags & MBX_FLAG_BOUNDARY) != 0;
    const bool logits = (p.flags & MBX_FLAG_LOGITS) != 0;
    const bool has_priors = !boundary;
    rcarve(&s, smem_raw, p.P, p.M, NWARPS, has_priors, !RS);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = p.P, M = p.M;
    const int Mp = rows_padded(M);
    const float half_alpha = __fdiv_rn(p.alpha, 2.0f);   // (alpha / 2.) in fp32, loss.py:35
    const double INF = CUDART_INF;
    unsigned status = 0;

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
    long long t_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long t_last = clock64();
    unsigned long long t_g0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_g0));
#endif

    unsigned invalid_mask = 0;   // columns of this thread beyond P
#pragma unroll
    for (int c = 0; c < C; ++c)
        if (tid + c * T >= P) invalid_mask |= 1u << c;

    const int n_work = static_cast<int>(gridDim.x);
    // Programmatic dependent launch (MBX_FLAG_PDL; the launcher sets the stream-serialization attribute):
    // this grid may start while the preceding kernel of the stream -- the previous training step -- is
    // still running; the next one may start as soon as every CTA of this grid is running.  The caller
    // promises that the INPUTS do not come from that preceding kernel; everything this grid shares with
    // it (outputs, workspace, all-reduce state) is only touched after griddepcontrol.wait, which returns
    // when the preceding grid has completed and its writes are visible.  So the load, the logs and the
    // whole assignment solve of step k+1 overlap the tail (epilogue, last-CTA reduction, completion) of
    // step k and the launch latency in between.  Without the attribute both instructions are no-ops.
    const bool pdl = (p.flags & MBX_FLAG_PDL) != 0;
    bool dep_done = !pdl;
    if (pdl) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        if (logits && p.conf_out) {   // (writes before the epilogue)
            asm volatile("griddepcontrol.wait;" ::: "memory");
            dep_done = true;
        }
    }

    // Image scheduling.  Static (image = CTA index, stride = resident CTAs) when every image has
    // its own CTA; otherwise DYNAMIC over the heavy-first order built by mbx_order_kernel: the
    // first wave takes positions 0..n_work-1, later positions are claimed from a global counter.
    // Thread 0 issues the claim when an image starts and reads it when the image is done, so the
    // L2 round trip of the atomic is off the critical path.
    const bool dyn = p.dynamic != 0;
    // The heavy-first order is computed by CTA 0 of this very kernel (no kernel in front of it), published
    // through a ready flag; the first wave either takes the images 0 .. order_first-1 in index order and
    // never waits, or (order_first == 0: large, skewed images, where a heavy-first first wave is worth 2 us)
    // waits for the order like everybody else.
    __shared__ int sh_sort[512];
    if (dyn && blockIdx.x == 0) {
        sort_images_heavy_first<T>(p, p.order_first, sh_sort, sh_sort + 256);
        __threadfence();
        if (tid == 0) st_release_gpu(p.oready, p.launch_id);
    }
    unsigned ar_seq_hint = 0xffffffffu;
    bool order_seen = !dyn;
    for (int q = static_cast<int>(blockIdx.x); q < p.B;) {
        int b = q;
        if (dyn && q >= p.order_first) {
            if (!order_seen) {
                while (ld_acquire_gpu(p.oready) != p.launch_id) {
                }
                order_seen = true;
            }
            b = __ldcg(p.order + (q - p.order_first));
        }
        unsigned claim = 0u;
        if (dyn && tid == 0) claim = atomicAdd(p.queue, 1u);
        const float4 *gg;
        int n = image_gt(p, b, gg);
        if (n < 0 || n > M) {
            status |= MBX_STATUS_BAD_NUM_GT;
            n = n < 0 ? 0 : M;
        }
        const size_t row0 = static_cast<size_t>(b) * P;
        // ---- per-column state in registers
        float4 loc[C];
        // lca: a FAST approximation of log(c) (2 instructions), good enough for the cheap cost form -- its
        // error is part of the margin (mbx_bound.h).  The exact numpy log of c (loss.py:21, ~45
        // instructions) is only needed where an exact cost is evaluated and for the matched priors'
        // loss term, so it is computed there, on demand: a few columns per row instead of all of them.
        float lca[C], l1[C], cf[C], wq[C];
        ColState<double, C, RS> cv, spc;     // dual v; shortest path cost of the current search
        ColState<short, C, RS> ptag, arow;   // visit index of the row that set spc; row the column had when scanned
        unsigned vnz = 0u;    // which of this thread's columns have a non-zero dual
        const float4 *gl = reinterpret_cast<const float4 *>(p.locations) + row0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = tid + c * T;
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
        // this thread's GT row (rows beyond the first T are loaded in the loop below), requested
        // together with the column loads: one memory round trip for the whole prologue
        float4 g_own = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid < n) g_own = gg[tid];
        if (!priors_ready) {
            mbar_wait(s.bar, 0);
            priors_ready = true;
        }
        unsigned lmax = 0u, tmax = 0u;   // bit images of max |coordinate| and max (|log c| + |log(1-c)|) of this thread's columns
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = tid + c * T;
            if (RS || j < P) {
                cv.set(c, j, s.cv, 0.0);
                arow.set(c, j, s.arow, static_cast<short>(-1));
            }
            if (j < P) {
                if (has_priors) {
                    const float4 q4 = s.priors[j];
                    loc[c].x = __fadd_rn(loc[c].x, q4.x);   // loss.py:71
                    loc[c].y = __fadd_rn(loc[c].y, q4.y);
                    loc[c].z = __fadd_rn(loc[c].z, q4.z);
                    loc[c].w = __fadd_rn(loc[c].w, q4.w);
                }
                if (logits) {
                    cf[c] = sigmoidf_(cf[c]);              // model.py:322
                    if (p.conf_out) p.conf_out[row0 + j] = cf[c];
                }
                const float ce = boundary ? cf[c] : __fadd_rn(cf[c], kEps32);   // loss.py:74
#ifdef MBX_EXP_EAGER_LC
                lca[c] = nplogf(ce);
#else
                lca[c] = __logf(ce);                                             // (see above; exact: exact_lc())
#endif
                float w = __fsub_rn(1.0f, ce);                                   // loss.py:22-24
                if (w > 1.0f) w = 1.0f;
                if (w <= 0.0f) w = kEps32;
                l1[c] = nplogf(w);                                               // loss.py:25
                if (n > 0) {
                    wq[c] = mbx_bound_w(loc[c].x, loc[c].y, loc[c].z, loc[c].w, half_alpha, lca[c], l1[c]);
                    // (uint order == float order on |x|; inf / NaN patterns come out on top and disable pruning)
                    const unsigned mx = max(max(__float_as_uint(fabsf(loc[c].x)), __float_as_uint(fabsf(loc[c].y))),
                                            max(__float_as_uint(fabsf(loc[c].z)), __float_as_uint(fabsf(loc[c].w))));
                    lmax = max(lmax, mx);
                    tmax = max(tmax, __float_as_uint(__fadd_rn(fabsf(lca[c]), fabsf(l1[c]))));
                } else {
                    wq[c] = 0.0f;
                }
            } else {
                lca[c] = 0.0f;           // a column that does not exist is never evaluated (invalid_mask)
                l1[c] = 0.0f;
                wq[c] = CUDART_INF_F;
            }
        }
        // exact log(c) of reference loss.py:21 for a confidence as loaded (bit-equal to numpy's float32 log)
#ifdef MBX_EXP_INLINE_LOG
        auto exact_lc = [&](float cfv) { return nplogf(boundary ? cfv : __fadd_rn(cfv, kEps32)); };
#else
        auto exact_lc = [&](float cfv) { return nplogf_cold(boundary ? cfv : __fadd_rn(cfv, kEps32)); };
#endif
        float m_w = CUDART_INF_F;   // margin of the cheap cost form for the columns of this warp
        if (n > 0) {
            lmax = __reduce_max_sync(0xffffffffu, lmax);
            tmax = __reduce_max_sync(0xffffffffu, tmax);
            m_w = mbx_bound_margin_col(__uint_as_float(lmax), __uint_as_float(tmax), half_alpha);
            if (lane == 0) s.mw[warp] = m_w;
        }
        for (int i = tid; i < n; i += T) {
            const float4 g = (i == tid) ? g_own : gg[i];
            float gpv[4], G, mg;
            mbx_bound_row(g.x, g.y, g.z, g.w, half_alpha, gpv, &G, &mg);
            s.gt[i] = g;
            s.gp[i] = make_float4(gpv[0], gpv[1], gpv[2], gpv[3]);
            s.rc[i] = make_float2(G, mg);
            s.u[i] = 0.0;
            s.col4row[i] = -1;
            s.rmin[i] = ~0ull;
            s.rmax[i] = 0ull;
            s.rcnt[i] = 0u;
        }
        for (int j = tid; j < P; j += T) {
            s.row4col[j] = -1;
            s.dirty[j] = 0;
        }
        if (tid == 0) s.ctl[1] = 0;
        block_sync<NWARPS>();
        MBX_T(0);   // prologue (load, logs)

        // ---- one shortest augmenting path per GT row (rows = GT, columns = priors)
        bool failed = false;
        bool ok = true;   // every cost entry evaluated so far is neither NaN nor -inf

        // ---- batched first Dijkstra step of EVERY row, assuming all column duals are zero.
        // Row i's first step is argmin_j (C(i,j) - v[j]); v is zero until an augmenting path
        // passes THROUGH a column, so all rows can be evaluated up front.
        // Pass 1 (no barrier, RB*C independent 4-FMA chains per thread): the cheap form of every
        // entry; per-warp minimum of each row -> rowpart.
        if (n > 0) {
            constexpr int RB = (C <= 3) ? 4 : 2;
            for (int i0 = 0; i0 < n; i0 += RB) {
                float4 gq[RB];
                float best[RB];
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    gq[r] = s.gp[(i0 + r < n) ? (i0 + r) : (n - 1)];
                    best[r] = CUDART_INF_F;
                }
#pragma unroll
                for (int c = 0; c < C; ++c) {
#pragma unroll
                    for (int r = 0; r < RB; ++r) best[r] = fminf(best[r], bound_a(loc[c], gq[r], wq[c]));
                }
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const float mn = warp_min_f32(best[r]);
                    if (lane == 0 && i0 + r < n) s.rowpart[warp * Mp + i0 + r] = mn;
                }
            }
            block_sync<NWARPS>();
            MBX_T(1);   // first step, pass 1 (cheap form of every entry)
            // Pass 2.  Upper bound of row i's true minimum: U = min over warps of (cheap minimum +
            // that warp's margin) [+ G_i + mg_i].  A column j of warp w can only be the minimum, or
            // tie for it, if its lower bound a - m_w [+ G_i - mg_i] does not exceed U, i.e.
            // a <= U + m_w + 2 mg_i (G_i cancels).  Each warp visits only the rows for which its own
            // cheap minimum passes that test, and evaluates cost32() only for the passing columns.
            // The exact evaluations are queued per LANE while the warp walks its rows (cheap, warp-uniform)
            // and run afterwards, all lanes of the warp side by side: one cost32() latency per warp instead
            // of one per visited row.
            int qe0 = -1, qe1 = -1;   // queued (row * 8 + c); a third candidate of one lane is evaluated on the spot
            auto eval_exact = [&](int entry) {
                const int row = entry >> 3, c = entry & 7;
                float4 l = loc[0];
                float cfv = cf[0], l1v = l1[0];
#pragma unroll
                for (int q2 = 1; q2 < C; ++q2) {
                    const bool hit = q2 == c;
                    l.x = hit ? loc[q2].x : l.x;
                    l.y = hit ? loc[q2].y : l.y;
                    l.z = hit ? loc[q2].z : l.z;
                    l.w = hit ? loc[q2].w : l.w;
                    cfv = hit ? cf[q2] : cfv;
                    l1v = hit ? l1[q2] : l1v;
                }
                const float c32 = cost32(l, s.gt[row], half_alpha, exact_lc(cfv), l1v);
                ok = ok && (c32 > -CUDART_INF_F);
                MBX_COUNT(9, 1);
                const unsigned long long k32 = ord32(c32);
                const unsigned long long col = static_cast<unsigned>(tid + c * T);
                if (atomicAdd(&s.rcnt[row], 1u) == 0u) {
                    s.rfirst[row] = (k32 << 32) | col;
                } else {
                    atomicMin(&s.rmin[row], (k32 << 32) | col);
                    atomicMax(&s.rmax[row], ((k32 ^ 0xffffffffull) << 32) | col);
                }
            };
            for (int k0 = 0; k0 < n; k0 += 32) {
                const int i = k0 + lane;
                float thr = 0.0f;
                bool flag = false;
                if (i < n) {
                    float U = CUDART_INF_F;
#pragma unroll
                    for (int w2 = 0; w2 < NWARPS; ++w2)
                        U = fminf(U, __fadd_ru(s.rowpart[w2 * Mp + i], s.mw[w2]));   // (a NaN sum bounds nothing: dropped)
                    const float mg = s.rc[i].y;
                    thr = __fadd_ru(__fadd_ru(U, m_w), __fadd_ru(mg, mg));
                    flag = !(s.rowpart[warp * Mp + i] > thr);
                }
                unsigned todo = __ballot_sync(0xffffffffu, flag);
                while (todo) {
                    const int bit = __ffs(todo) - 1;
                    todo &= todo - 1u;
                    const int row = k0 + bit;
                    const float t = __shfl_sync(0xffffffffu, thr, bit);
                    const float4 gq = s.gp[row];
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        if ((invalid_mask >> c) & 1u) continue;
                        if (bound_a(loc[c], gq, wq[c]) > t) continue;   // (a NaN bound fails the test: evaluated)
                        const int entry = row * 8 + c;
                        if (qe0 < 0)
                            qe0 = entry;
                        else if (qe1 < 0)
                            qe1 = entry;
                        else
                            eval_exact(entry);
                    }
                }
            }
            if (qe0 >= 0) eval_exact(qe0);
            if (qe1 >= 0) eval_exact(qe1);
            block_sync<NWARPS>();
            MBX_T(8);   // first step, pass 2 (exact costs of the candidates)
        }

        // ---- rows in order.  Warp 0 disposes of every row whose precomputed first step is
        // decisive (unique minimum at a column whose dual is still zero and which is unassigned:
        // that column is the sink, the path is the single edge, no dual changes besides
        // u[row] = min).  A dual can only make its column MORE expensive (v <= 0, enforced below),
        // so a zero-dual minimum stays the true minimum.  Any other row is solved by the whole CTA
        // with the general shortest-augmenting-path search below.
        int cur = 0;
        for (;;) {
            if (warp == 0) {
                // Up to 32 consecutive rows per round: lane r takes row cur+r.  A row is decisive when
                // its first-step minimum is unique, finite, at a column with zero dual that is
                // unassigned AND not wanted by an earlier row of the same round.  The round commits
                // the rows before the first non-decisive one.
                const bool fast_off = s.ctl[1] != 0;
                while (cur < n && !fast_off) {
                    const int row = cur + lane;
                    bool good = false;
                    unsigned col = kColNone, key = 0u;
                    if (row < n) {
                        bool unique;
                        first_step_result(s, row, key, col, unique);
                        good = unique && key < kOrdInf32;   // one column at the minimum, finite
                        if (good) good = !s.dirty[col] && s.row4col[col] < 0;
                    }
                    // an earlier lane of this round wants the same column -> this row conflicts
                    const unsigned same = __match_any_sync(0xffffffffu, col);
                    if (good && (same & ((1u << lane) - 1u))) good = false;
                    const unsigned bad = __ballot_sync(0xffffffffu, !good);   // rows >= n are "bad" too
                    const int nfast = bad ? (__ffs(bad) - 1) : 32;
                    if (lane < nfast) {
                        s.row4col[col] = static_cast<short>(row);
                        s.col4row[row] = static_cast<int>(col);
                        s.u[row] = static_cast<double>(unord32(key));
                    }
                    cur += nfast;
                    __syncwarp();
                    if (nfast < 32) break;
                }
                if (lane == 0) s.ctl[0] = cur;
            }
            block_sync<NWARPS>();
            cur = s.ctl[0];
            MBX_T(2);   // sequential fast rows
            if (cur >= n) break;
            // assigned bits of this thread's columns (warp 0 assigned sinks on its own)
            unsigned asg = 0u;
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (!((invalid_mask >> c) & 1u) && s.row4col[tid + c * T] >= 0) asg |= 1u << c;
            int i = cur, R = 0;
            double min_val = 0.0, ui = 0.0;
            unsigned scmask = invalid_mask;   // columns already scanned (or non-existent)
            float UB = CUDART_INF_F;          // upper bound (rounded up) of this search's final path cost
            const bool vpos = s.ctl[1] != 0;  // some dual went positive: a dirty column is never pruned
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (!((invalid_mask >> c) & 1u)) spc.set(c, tid + c * T, s.spc, INF);
            // The first Dijkstra step of this row is already known when its precomputed first-step
            // minimum is unique and sits at a column whose dual is still zero: duals only raise
            // reduced costs (v <= 0), so that column is the strict minimum of C(cur, j) - v[j] over
            // all j, exactly what scanning row `cur` would select (the same argument as for the
            // decisive rows above; the column is assigned, or the row would have been decisive).
            // The block arg-min and the barrier of step 0 are skipped; the path costs step 0 would
            // have left behind are formed together with step 1's (`pend0`).
            bool pend0 = false;
            {
                unsigned key0, col0;
                bool unique0;
                first_step_result(s, cur, key0, col0, unique0);
                if (!vpos && unique0 && key0 < kOrdInf32 && !s.dirty[col0]) {
                    const int r4c0 = s.row4col[col0];
                    if (r4c0 >= 0) {
                        pend0 = true;
                        min_val = static_cast<double>(unord32(key0));
                        const int cstar0 = static_cast<int>(col0) / T;
                        if (static_cast<int>(col0) - cstar0 * T == tid) {   // the column's owner logs the removal
#pragma unroll
                            for (int c = 0; c < C; ++c)
                                if (c == cstar0) {
                                    arow.set(c, static_cast<int>(col0), s.arow, static_cast<short>(r4c0));
                                    spc.set(c, static_cast<int>(col0), s.spc, min_val);   // C(cur, col0) - 0: what step 0 stores
                                    ptag.set(c, static_cast<int>(col0), s.pmv, static_cast<short>(0));
                                }
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
            bool first_scan = true;
            for (;;) {
                // ---- scan: path costs through row i (after a skipped step 0: through row cur first)
#pragma unroll 1
                for (int pass = pend0 ? 0 : 1; pass < 2; ++pass) {
                    const int row = pass ? i : cur;
                    const double mv = pass ? min_val : 0.0, uu = pass ? ui : 0.0;
                    const short tag = static_cast<short>(pass ? R : 0);
                    const float4 gq = s.gp[row];
                    const float4 g = s.gt[row];
                    const float2 rcst = s.rc[row];
                    float a[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) a[c] = bound_a(loc[c], gq, wq[c]);
                    // path cost through this row: r = ((mv + C) - uu) - v in fp64, v <= 0.  With base = mv - uu and
                    // S = |mv| + |uu| (+ |UB|), the fp64 chain differs from base + C by at most 2^-52 (S + |C|): the
                    // 2^-40 S terms below cover it (for |C| > 2^11 S the comparisons hold trivially).
                    const float base_up = __double2float_ru(__dsub_rn(mv, uu));
                    const float base_dn = __double2float_rd(__dsub_rn(mv, uu));
                    const float S = __fadd_ru(__double2float_ru(fabs(mv)), __double2float_ru(fabs(uu)));
                    const float krow = __fadd_ru(__fadd_ru(rcst.y, m_w), -rcst.x);   // mg + m_w - G (rounded up)
                    if (first_scan) {
                        // no block-wide bound yet: the cheapest UNASSIGNED column of this warp under the cheap
                        // form bounds the final path cost (the search ends at an unassigned column whose path
                        // cost is at most that of any unassigned column it has seen; their duals are zero)
                        float am = CUDART_INF_F;
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            if (!(((asg | scmask) >> c) & 1u)) am = fminf(am, a[c]);
                        am = warp_min_f32(am);
                        const float ubc = __fadd_ru(__fadd_ru(am, rcst.x), __fadd_ru(rcst.y, m_w));   // >= its exact cost
                        float ub = __fadd_ru(base_up, ubc);
                        ub = __fmaf_ru(__fadd_ru(S, fabsf(ubc)), 9.094947017729282e-13f, ub);          // + 2^-40 (...)
                        UB = fminf(UB, fminf(ub, CUDART_INF_F));   // (NaN -> +inf)
                        first_scan = false;
                    }
                    // a column whose exact path cost certainly exceeds UB is skipped: C > UB - base (+ slack) is
                    // implied by a > thr (mbx_bound.h; every operation rounds up)
                    float thr = __fadd_ru(__fsub_ru(UB, base_dn), krow);
                    thr = __fmaf_ru(__fadd_ru(S, fabsf(UB)), 9.094947017729282e-13f, thr);
                    unsigned cand = 0u;
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        if (!((scmask >> c) & 1u) && !(a[c] > thr)) cand |= 1u << c;
                    if (vpos) cand |= vnz & ~scmask;
                    // exact costs of the candidates: every lane takes its lowest pending column, all lanes side by side
                    while (__any_sync(0xffffffffu, cand != 0u)) {
                        if (cand != 0u) {
                            const int c = __ffs(cand) - 1;
                            cand &= cand - 1u;
                            float4 l = loc[0];
                            float cfv = cf[0], l1v = l1[0];
#pragma unroll
                            for (int q2 = 1; q2 < C; ++q2) {
                                const bool hit = q2 == c;
                                l.x = hit ? loc[q2].x : l.x;
                                l.y = hit ? loc[q2].y : l.y;
                                l.z = hit ? loc[q2].z : l.z;
                                l.w = hit ? loc[q2].w : l.w;
                                cfv = hit ? cf[q2] : cfv;
                                l1v = hit ? l1[q2] : l1v;
                            }
                            const float c32 = cost32(l, g, half_alpha, exact_lc(cfv), l1v);
                            ok = ok && (c32 > -CUDART_INF_F);
                            MBX_COUNT(9, 1);
                            const int jc = tid + c * T;
                            const double r =
                                __dsub_rn(__dsub_rn(__dadd_rn(mv, static_cast<double>(c32)), uu), cv.get(c, jc, s.cv));
                            if (r < spc.get(c, jc, s.spc)) {
                                spc.set(c, jc, s.spc, r);
                                ptag.set(c, jc, s.pmv, tag);
                            }
                        }
                    }
                }
                pend0 = false;
                // ---- thread-local minimum over the live columns; cheapest unassigned one for UB
                double best = INF;
                unsigned bj = kPayNone, tie = 0u;
                float ubl = CUDART_INF_F;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const bool live = !((scmask >> c) & 1u);
                    double sp = INF;
                    if (live) sp = spc.get(c, tid + c * T, s.spc);
                    const bool lt = sp < best;
                    const unsigned eq = sp == best ? 1u : 0u;
                    best = lt ? sp : best;
                    bj = lt ? ((static_cast<unsigned>(tid + c * T) << 1) | ((asg >> c) & 1u)) : bj;
                    tie = lt ? 0u : (tie | eq);
                    if (live && !((asg >> c) & 1u)) ubl = fminf(ubl, __double2float_ru(sp));
                }
                const unsigned long long key = (bj == kPayNone) ? ~0ull : ord64(best);
                MBX_T(3);   // general scan
                // ---- block-wide arg-min of (path cost, column); exact ties flagged
                unsigned hi = static_cast<unsigned>(key >> 32), lo = static_cast<unsigned>(key);
                unsigned pay = bj;
                bool tieb = tie != 0u;
                warp_argmin(hi, lo, pay, tieb);
                ubl = warp_min_f32(ubl);
                if (NWARPS > 1) {
                    if (lane == 0)
                        s.part[pbuf * NWARPS + warp] =
                            make_int4(static_cast<int>(hi), static_cast<int>(lo),
                                      static_cast<int>(pay | (tieb ? 0x80000000u : 0u)), __float_as_int(ubl));
                    block_sync<NWARPS>();
                    int4 e = make_int4(-1, -1, static_cast<int>(kPayNone), __float_as_int(CUDART_INF_F));
                    if (lane < NWARPS) e = s.part[pbuf * NWARPS + lane];
                    hi = static_cast<unsigned>(e.x);
                    lo = static_cast<unsigned>(e.y);
                    pay = static_cast<unsigned>(e.z) & 0x7fffffffu;
                    tieb = (static_cast<unsigned>(e.z) >> 31) != 0u;
                    warp_argmin(hi, lo, pay, tieb);
                    ubl = warp_min_f32(__int_as_float(e.w));
                    pbuf ^= 1;
                }
                UB = fminf(UB, ubl);
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
                        const double sp = spc.get(c, tid + c * T, s.spc);
                        if (!(sp == min_val)) continue;
                        const int j = tid + c * T;
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
                    if (NWARPS > 1) {
                        if (lane == 0) s.pk[warp] = k;
                        block_sync<NWARPS>();
                        k = s.pk[0];
#pragma unroll
                        for (int w = 1; w < NWARPS; ++w) k = s.pk[w] < k ? s.pk[w] : k;
                        block_sync<NWARPS>();
                    }
                    jstar = static_cast<int>((k & 0xffffffffu) >> 1);
                    is_sink = !(k & 1u);
                }
                const int cstar = jstar / T;
                const bool owner = (jstar - cstar * T) == tid;
                if (is_sink) {
                    if (owner) {
                        // ---- the sink's owner augments along the path back to row `cur`.  Every log
                        // entry it reads was written before an earlier barrier; the sink itself needs
                        // no log entry (nothing is scanned after it).
                        int pmv = 0;
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            if (c == cstar) pmv = ptag.get(c, jstar, s.pmv);
                        scmask |= 1u << cstar;
                        asg |= 1u << cstar;
                        s.u[cur] = min_val;          // u[cur] was 0: 0 + min_val
                        int col = jstar, m = pmv;
                        for (;;) {
                            const int row = (m == 0) ? cur : s.visit[m];
                            s.row4col[col] = static_cast<short>(row);
                            s.col4row[row] = col;
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
                            pmv = ptag.get(c, jstar, s.pmv);
                            arow.set(c, jstar, s.arow, static_cast<short>(r4c_star));
                        }
                    scmask |= 1u << cstar;
                    s.rm_col[R] = jstar;
                    s.rm_idx[R] = replay_pos(jstar, R, P, s.rm_idx);
                    s.rm_pm[R] = pmv;
                    s.visit[R + 1] = r4c_star;
                }
                ++R;
                i = r4c_star;
                ui = s.u[i];
                if (NWARPS == 1) __syncwarp();
            }
            if (failed) break;
            MBX_T(5);   // selection, log, walk
            // ---- dual update: v (owner), u of the visited rows (shared, distinct rows).
            // Only columns scanned BEFORE the sink move (the sink's own delta is 0).
            if (R > 1) {
                const unsigned scanned = scmask & ~invalid_mask;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int j = tid + c * T;
                    if (!((scanned >> c) & 1u)) continue;
                    const int ar = arow.get(c, j, s.arow);
                    if (ar >= 0) {
                        const double sp = spc.get(c, j, s.spc);
                        const double delta = __dsub_rn(min_val, sp);
                        const double vn = __dsub_rn(cv.get(c, j, s.cv), delta);
                        cv.set(c, j, s.cv, vn);
                        s.u[ar] = __dadd_rn(s.u[ar], delta);
                        if (vn != 0.0) {
                            vnz |= 1u << c;
                            s.dirty[j] = static_cast<unsigned char>(1);
                            // the precomputed first steps and the pruning rely on v <= 0; fp rounding could in
                            // principle break that by an ulp: then every later row goes general, unpruned there
                            if (vn > 0.0) s.ctl[1] = 1;
                        }
                        arow.set(c, j, s.arow, static_cast<short>(-1));
                    }
                }
            }
            ++cur;
            block_sync<NWARPS>();   // walk, duals and dirty marks visible to warp 0
        }
        MBX_T(6);   // dual update
        if (!ok) status |= MBX_STATUS_INVALID_COST;
        block_sync<NWARPS>();   // the last walk's row4col / col4row are visible below

        // ---- epilogue, part 1 (no global write yet): loss terms and gradients of this thread's columns in
        // registers, block reduction of the loss sums -- all of it still overlaps the preceding grid
        double acc_sq = 0.0, acc_conf = 0.0;
        int n_match = 0;
        float4 dlv[C];
        float dcv[C];
        int rv[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = tid + c * T;
            dlv[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            dcv[c] = 0.f;
            rv[c] = -1;
            if (j >= P) continue;
            const int r = s.row4col[j];
            rv[c] = r;
            n_match += r >= 0;
            const float ce = boundary ? cf[c] : __fadd_rn(cf[c], kEps32);
            // unmatched prior (almost all of them), unconditionally: -log((1 - c) + eps) and its derivative
            // (__frcp_rn is the correctly rounded reciprocal: the same bits as the IEEE division 1 / x)
            const float one_m = __fsub_rn(1.0f, ce);
            const float arg = __fadd_rn(one_m, kEps32);   // loss.py:101
            float vcl = one_m;
            if (vcl > 1.0f) vcl = 1.0f;
            if (vcl <= 0.0f) vcl = kEps32;
            float la = l1[c];
#ifdef MBX_EXP_INLINE_LOG
            if (arg != vcl) la = nplogf(arg);             // (only for saturated confidences)
#else
            if (arg != vcl) la = nplogf_cold(arg);        // (only for saturated confidences)
#endif
            float4 dl = make_float4(0.f, 0.f, 0.f, 0.f);
            float dc = __frcp_rn(arg);
            if (r >= 0) {                                 // matched: location term, -log(c), their derivatives
                const float4 g = s.gt[r];
                const float d0 = __fsub_rn(loc[c].x, g.x), d1 = __fsub_rn(loc[c].y, g.y),
                            d2 = __fsub_rn(loc[c].z, g.z), d3 = __fsub_rn(loc[c].w, g.w);
                acc_sq += static_cast<double>(__fmul_rn(d0, d0));
                acc_sq += static_cast<double>(__fmul_rn(d1, d1));
                acc_sq += static_cast<double>(__fmul_rn(d2, d2));
                acc_sq += static_cast<double>(__fmul_rn(d3, d3));
                dl = make_float4(__fmul_rn(p.alpha, d0), __fmul_rn(p.alpha, d1), __fmul_rn(p.alpha, d2),
                                 __fmul_rn(p.alpha, d3));
                la = exact_lc(cf[c]);
                dc = -__frcp_rn(ce);
            }
            acc_conf -= static_cast<double>(la);
            if (logits) dc = __fmul_rn(dc, __fmul_rn(cf[c], __fsub_rn(1.0f, cf[c])));
            dlv[c] = dl;
            dcv[c] = dc;
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
        // ---- epilogue, part 2: the stores
        if (!dep_done) {   // first global write of this CTA: the preceding grid must be complete (see above)
            asm volatile("griddepcontrol.wait;" ::: "memory");
            dep_done = true;
        }
        // (fused all-reduce: the step counter is final once the preceding grid is complete; whoever turns out
        // to be the last CTA then knows the step to collect before it has reduced anything)
        if (p.ar_world > 1 && tid == 0) ar_seq_hint = __ldcg(p.ar_seq);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = tid + c * T;
            if (j >= P) continue;
#ifndef MBX_PHASE_TIMING   // (timing builds use the mask buffer for the cycle counters)
            if (p.mask) p.mask[row0 + j] = rv[c] >= 0 ? 1 : 0;
#endif
            if (p.gt_idx) p.gt_idx[row0 + j] = rv[c];
            if (nheads > 1) {   // gradients in the per-head layouts too
                int hh;
                const size_t e = head_elem(sh_heads, nheads, j, b, hh);
                if (sh_heads[hh].dloc) st_stream_f4(reinterpret_cast<float4 *>(sh_heads[hh].dloc) + e, dlv[c]);
                if (sh_heads[hh].dconf) sh_heads[hh].dconf[e] = dcv[c];
            } else {
                if (p.d_loc) st_stream_f4(reinterpret_cast<float4 *>(p.d_loc) + row0 + j, dlv[c]);
                if (p.d_conf) p.d_conf[row0 + j] = dcv[c];
            }
        }
        if (p.stacked && !failed) {
            const int off = p.stk_offsets[b];
            for (int i = tid; i < n; i += T) {
                const int pi = s.col4row[i];
                int rank = 0;
                for (int q2 = 0; q2 < n; ++q2) rank += s.col4row[q2] < pi;
                reinterpret_cast<float4 *>(p.stacked)[off + rank] = s.gt[i];
            }
        }
        if (tid == 0) {
            double a = 0.0, cc = 0.0;
            int m = 0;
            for (int w = 0; w < NWARPS; ++w) {
                a += s.red[w];
                cc += s.red[NWARPS + w];
                m += s.ri[w];
            }
            p.partials[2 * b] = a;
            p.partials[2 * b + 1] = cc;
            p.img_matched[b] = m;
        }
        if (dyn && tid == 0) s.ctl[2] = n_work + static_cast<int>(claim);
        block_sync<NWARPS>();   // shared state is reused by the next image
        q = dyn ? s.ctl[2] : q + n_work;
    }

    if (!dep_done) asm volatile("griddepcontrol.wait;" ::: "memory");   // (a CTA that solved no image)
    if (status) {
        atomicOr(p.status, status);
        __threadfence();
    }
#ifdef MBX_PHASE_TIMING
    MBX_T(7);   // epilogue
    t_acc[9] = __reduce_add_sync(0xffffffffu, static_cast<int>(t_acc[9]));   // exact cost evaluations of the warp
    if (lane == 0 && p.mask) {
        long long *dbg = reinterpret_cast<long long *>(p.mask) +
                         (static_cast<size_t>(blockIdx.x) * NWARPS + warp) * kTimingSlots;
        for (int k = 0; k < 10; ++k) dbg[k] = t_acc[k];
        unsigned long long t_g1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_g1));
        dbg[10] = static_cast<long long>(t_g0);
        dbg[11] = static_cast<long long>(t_g1);
    }
#endif

    // ---- last CTA to finish reduces the per-image partials in a fixed order.  Thread 0 wrote this CTA's
    // partials itself; its acq_rel ticket atomic publishes them (and, through the barrier, whatever the other
    // threads' status atomics did) and acquires the other CTAs' -- no block-wide __threadfence, which would
    // wait for every thread's gradient stores to drain, on the one chain that is not overlapped.
    __shared__ bool is_last;
    block_sync<NWARPS>();
    if (tid == 0) {
        unsigned t;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(p.ticket) : "memory");
        is_last = (t == gridDim.x - 1);
    }
    block_sync<NWARPS>();
    if (!is_last) return;
    // status word and previous launch sequence number: loaded together with the partials
    const TailPrefetch pre = tail_prefetch(p);
    CollectPrefetch cpf;
    cpf.step = 0xffffffffu;
    if (warp == 0 && p.ar_world > 1) {   // the collect's words (own table), requested in the same round as the partials
        const unsigned sh = __shfl_sync(0xffffffffu, ar_seq_hint, 0);
        const unsigned lag = (p.flags & MBX_FLAG_AR_DEFERRED) ? ar_lag(p) : 0u;
        cpf = ar_collect_prefetch(p.ar_peer, p.ar_world, p.ar_rank, (sh != 0xffffffffu && sh >= lag) ? sh - lag : 0xffffffffu);
    }
    double a = 0.0, cc = 0.0, md = 0.0;
    for (int b = tid; b < p.B; b += T) {
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
        double A = 0.0, Cc = 0.0, Mt = 0.0;
        for (int w = 0; w < NWARPS; ++w) {
            A += s.red[w];
            Cc += s.red[NWARPS + w];
            Mt += s.red[2 * NWARPS + w];
        }
        finalize_losses(p, A, Cc, Mt, pre, &cpf);   // warp-cooperative (lane r posts to peer r)
    }
}

namespace {

struct KernelInfo {
    int dev = -1;
    size_t configured_smem = 0;
    int occ = 0;
    size_t occ_smem = 0;
};

template <int NWARPS, int C>
int launch_one(const MatchParams &p, cudaStream_t st) {
    static thread_local KernelInfo info;
    auto kern = mbx_match_loss_reg_kernel<NWARPS, C>;
    const size_t smem = rcarve(nullptr, nullptr, p.P, p.M, NWARPS, !(p.flags & MBX_FLAG_BOUNDARY), C > 3);
    if (smem > static_cast<size_t>(max_smem_optin())) return MBX_E_TOO_LARGE;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != info.dev) {   // function attributes and occupancy are per device
        info = KernelInfo();
        info.dev = dev;
    }
    if (smem > info.configured_smem) {
        if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    static_cast<int>(smem)),
                               "cudaFuncSetAttribute(match_reg)"))
            return e;
        // (the whole L1/shared array as shared memory: the kernel streams its global data past L1 anyway)
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        info.configured_smem = smem;
        info.occ = 0;
    }
    if (info.occ == 0 || info.occ_smem != smem) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&info.occ, kern, NWARPS * 32, smem);
        if (info.occ < 1) info.occ = 1;
        info.occ *= sm_count();        // resident CTAs on the whole device
        info.occ_smem = smem;
    }
    int units = info.occ;
    if (units > p.B) units = p.B;
    MatchParams pp = p;
    if (p.B > units && !(p.flags & MBX_FLAG_STATIC)) {
        // more images than resident CTAs: heavy-first order (computed by CTA 0 of the kernel) + dynamic
        // scheduling.  Consecutive launches on one workspace alternate between two scheduler slots.
        const unsigned id = next_launch_id(p.ticket);   // (one counter per workspace, shared by all instantiations)
        const unsigned slot = id & 1u;
        pp.dynamic = 1;
        pp.launch_id = id;
        pp.order = p.order_base + static_cast<size_t>(slot) * p.B;
        pp.queue = p.sched_base + slot;
        pp.oready = p.sched_base + 2 + slot;
        pp.order_first = (p.M >= 128) ? 0 : units;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(units);
    cfg.blockDim = dim3(NWARPS * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pp.flags & MBX_FLAG_PDL) ? 1 : 0;
    return check_cuda(cudaLaunchKernelEx(&cfg, kern, pp), "launch mbx_match_loss_reg_kernel");
}

}  // namespace

template <int NWARPS>
int launch_cols(const MatchParams &p, int cols, cudaStream_t st) {
    switch (cols) {
        case 1: return launch_one<NWARPS, 1>(p, st);
        case 2: return launch_one<NWARPS, 2>(p, st);
        case 3: return launch_one<NWARPS, 3>(p, st);
        case 4: return launch_one<NWARPS, 4>(p, st);
        case 5: return launch_one<NWARPS, 5>(p, st);
        case 6: return launch_one<NWARPS, 6>(p, st);
        case 7:
        case 8: return launch_one<NWARPS, 8>(p, st);
        default: return MBX_E_TOO_LARGE;
    }
}

}  // namespace mbx
