// multibox_b200 -- detection post-processing, one CTA per image / patch
// (persistent grid), sm_100a.
//
// What it replaces (reference = gvanhorn38/multibox):
//   detect.py:412-413  decode (offset + prior) and clip to [0,1]       -> load phase
//   detect.py:74-104   filter_proposals (per-box python loop)          -> load phase (predicate)
//   detect.py:423-427  argsort(conf)[::-1][:max_to_keep]               -> 64-bit key sort in shared
//                      memory; key = (orderable(conf) << 32 | prior index), descending, which is the
//                      order numpy's stable argsort + reversal yields (ties: descending index)
//   (extension)        greedy NMS on the kept boxes, bitmask formulation: one warp ballot per
//                      (row, 32-column word) builds the suppression matrix in shared memory, one
//                      warp sweeps it in score order
//   detect.py:106-131  convert_proposals (float64 scale/offset/flip)   -> store phase
//   model.py:322       sigmoid (MBX_FLAG_LOGITS)                       -> load phase
//   eval.py:146-167    decode / clip / full sort / top-100             -> same kernel, no filter
#include <math_constants.h>

#include <type_traits>

#include "mbx_common.cuh"

// Optional phase timing (profiles/phase_timing.py --detect builds with -DMBX_PHASE_TIMING).
#ifdef MBX_PHASE_TIMING
#define MBX_DT(k)                                        \
    do {                                                 \
        const long long t_now__ = clock64();             \
        t_acc[k] += t_now__ - t_last;                    \
        t_last = t_now__;                                \
    } while (0)
#else
#define MBX_DT(k) \
    do {          \
    } while (0)
#endif

namespace mbx {

struct DetectParams {
    const float *locations, *confidences, *priors, *restrictions;
    const int32_t *max_to_keep, *offsets, *patch_dims, *image_dims, *is_flipped;
    int B, P, k_max, n2;
    int nheads;                       // > 1: inputs come per head (model.py:295-320 never materialised)
    HeadTab heads[MBX_MAX_HEADS];
    float nms_iou;
    unsigned flags;
    double *out_boxes;
    float *out_patch_boxes, *out_scores;
    int32_t *out_idx, *out_count;
};

// monotone float32 -> uint32 map (larger float => larger uint)
__device__ __forceinline__ uint32_t orderable(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__device__ __forceinline__ float clip01(float x) { return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); }

// fp32 IoU test exactly as oracle/np_oracle.greedy_nms (and torchvision's CPU nms) decides it:
//   fl(inter / ((area_a + area_b) - inter)) > thr.
// iou_fast decides WITHOUT a division: with den > 0, x = inter/den and t = fl(thr*den),
//   inter > t*(1+2^-20)  =>  x > thr*(1+2^-21) > nextafter(thr)  =>  fl(x) > thr     (rounding is monotone)
//   inter < t*(1-2^-20)  =>  x < thr                             =>  fl(x) <= thr
// so only pairs with |inter - t| <= 2^-20*t (or a denominator that is not a comfortably normal
// positive number: tiny, inf) are `near`; those (rare) pairs are re-decided with the
// IEEE division (iou_exact), and the final decision is always the exact one.  The caller must
// route EVERY pair through iou_exact when thr itself is outside [2^-20, 2^20] (thr_ok below).
// The claim is also checked numerically on the host by tests/test_nms_decision_math.py.
__device__ __forceinline__ bool iou_fast(float4 a, float area_a, float4 b, float area_b, float thr, bool &near) {
    const float w = fmaxf(0.0f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
    const float h = fmaxf(0.0f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
    const float inter = __fmul_rn(w, h);
    const float den = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    const float t = __fmul_rn(thr, den);
    const float d = __fsub_rn(inter, t);
    // den <= 0 needs no second look either (inter >= 0, thr > 0): den < 0 (a box with x2 < x1 -- the
    // reference applies no validity fix-up, detect.py:412-413) gives a quotient <= 0, never > thr;
    // den == 0 gives t = 0, d = inter, and inter/0 > thr <=> inter > 0.
    // (every operand is computed before the predicates are combined, so that the combination is
    // pure predicate logic: a short-circuit around arithmetic makes the compiler branch per pair)
    const float tolt = __fmul_rn(9.5367431640625e-07f, t);
    const bool pos = den > 0.0f, normal = den >= 1e-20f, clear = fabsf(d) > tolt, nonneg = den >= 0.0f, gt = d > 0.0f;
    near = pos && !(normal && clear);
    return gt && nonneg;
}
__device__ __noinline__ bool iou_exact(float4 a, float area_a, float4 b, float area_b, float thr) {
    const float w = fmaxf(0.0f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
    const float h = fmaxf(0.0f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
    const float inter = __fmul_rn(w, h);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter)) > thr;
}

constexpr int kBins = 1024;   // confidence histogram used to pre-select the top max_to_keep
__device__ __forceinline__ int conf_bin(float c) {
    // monotone non-decreasing in the sort order of c (any float); NaN sorts first like numpy's
    if (c != c) return kBins - 1;
    const float t = fminf(fmaxf(c * static_cast<float>(kBins), 0.0f), static_cast<float>(kBins - 1));
    return static_cast<int>(t);
}

struct DSmem {
    float4 *priors, *box, *sbox;
    unsigned long long *keys, *ckeys;
    int *hist;
    float *sarea;
    uint32_t *diag, *remw, *keepw;
    int *klist;
    int *keeppre, *cnt;
    uint64_t *bar;
};

__host__ __device__ inline size_t dalign(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline size_t dcarve(DSmem *s, unsigned char *base, int P, int n2, int k_max, bool nms,
                                         bool has_priors) {
    const int W = (k_max + 31) / 32;
    size_t o = 0;
    auto take = [&](size_t bytes, size_t al) {
        o = dalign(o, al);
        size_t r = o;
        o += bytes;
        return r;
    };
    size_t o_pri = take(has_priors ? sizeof(float4) * P : 0, 16);
    size_t o_box = take(sizeof(float4) * P, 16);
    size_t o_sbox = take(nms ? sizeof(float4) * (W * 32) : 0, 16);   // padded to whole 32-row blocks
    size_t o_keys = take(sizeof(unsigned long long) * n2, 8);
    size_t o_ckeys = take(sizeof(unsigned long long) * n2, 8);
    size_t o_bar = take(8, 8);
    size_t o_hist = take(sizeof(int) * kBins, 4);
    size_t o_area = take(nms ? sizeof(float) * (W * 32) : 0, 4);
    size_t o_nm = take(nms ? sizeof(uint32_t) * (W * 32) : 0, 4);     // diag[i]: whom box i suppresses inside its chunk
    size_t o_tc = take(sizeof(uint32_t) * 32, 4);                      // remw[w]: boxes of chunk w removed by earlier chunks
    size_t o_kw = take(sizeof(uint32_t) * 32, 4);
    size_t o_kp = take(sizeof(int) * 33, 4);
    size_t o_kl = take(sizeof(int) * 64, 4);       // kept rows of a chunk, double buffered
    size_t o_cnt = take(sizeof(int) * 8, 4);
    if (s) {
        s->priors = reinterpret_cast<float4 *>(base + o_pri);
        s->box = reinterpret_cast<float4 *>(base + o_box);
        s->sbox = reinterpret_cast<float4 *>(base + o_sbox);
        s->keys = reinterpret_cast<unsigned long long *>(base + o_keys);
        s->ckeys = reinterpret_cast<unsigned long long *>(base + o_ckeys);
        s->hist = reinterpret_cast<int *>(base + o_hist);
        s->bar = reinterpret_cast<uint64_t *>(base + o_bar);
        s->sarea = reinterpret_cast<float *>(base + o_area);
        s->diag = reinterpret_cast<uint32_t *>(base + o_nm);
        s->keepw = reinterpret_cast<uint32_t *>(base + o_kw);
        s->remw = reinterpret_cast<uint32_t *>(base + o_tc);
        s->keeppre = reinterpret_cast<int *>(base + o_kp);
        s->klist = reinterpret_cast<int *>(base + o_kl);
        s->cnt = reinterpret_cast<int *>(base + o_cnt);
    }
    return dalign(o, 16);
}

// Descending bitonic sort of N = T*R 64-bit keys held R per thread (thread `tid` owns elements
// tid*R .. tid*R+R-1).  Compare-exchange distances below R run in registers, below 32*R with warp
// shuffles; only the distances that cross warps go through shared memory (9 barriers for
// N = 1024 on 256 threads instead of the 55 of an all-shared-memory sort).  On return the
// sorted keys are in `sm[0..N)` (and in the registers).
template <int T, int R>
__device__ __forceinline__ void block_sort_desc(unsigned long long (&k)[R], unsigned long long *sm, int tid) {
    constexpr int N = T * R;
#pragma unroll
    for (int kk = 2; kk <= N; kk <<= 1) {
        bool in_smem = false;
#pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1) {
            if (j >= 32 * R) {
                if (!in_smem) {
#pragma unroll
                    for (int r = 0; r < R; ++r) sm[tid * R + r] = k[r];
                    __syncthreads();
                    in_smem = true;
                }
                for (int t = tid; t < (N >> 1); t += T) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int ixj = i | j;
                    const unsigned long long a = sm[i], c = sm[ixj];
                    const bool desc = (i & kk) == 0;
                    if ((a < c) == desc) {
                        sm[i] = c;
                        sm[ixj] = a;
                    }
                }
                __syncthreads();
            } else {
                if (in_smem) {
#pragma unroll
                    for (int r = 0; r < R; ++r) k[r] = sm[tid * R + r];
                    in_smem = false;
                }
                if (j >= R) {
                    const int lm = j / R;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const unsigned long long o = __shfl_xor_sync(0xffffffffu, k[r], lm);
                        const int e = tid * R + r;
                        const bool keep_max = (((e & kk) == 0) == ((e & j) == 0));
                        const unsigned long long mx = k[r] > o ? k[r] : o, mn = k[r] > o ? o : k[r];
                        k[r] = keep_max ? mx : mn;
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if ((r & j) == 0) {
                            const int e = tid * R + r;
                            const bool desc = (e & kk) == 0;
                            const unsigned long long a = k[r], c = k[r | j];
                            const unsigned long long mx = a > c ? a : c, mn = a > c ? c : a;
                            k[r] = desc ? mx : mn;
                            k[r | j] = desc ? mn : mx;
                        }
                    }
                }
            }
        }
    }
    __syncthreads();   // everyone is done reading sm from the last cross-warp phase
#pragma unroll
    for (int r = 0; r < R; ++r) sm[tid * R + r] = k[r];
    __syncthreads();
}

template <int T, int R>
__device__ __forceinline__ void sort_keys_in_smem(unsigned long long *sm, int tid) {
    unsigned long long k[R];
#pragma unroll
    for (int r = 0; r < R; ++r) k[r] = sm[tid * R + r];
    block_sort_desc<T, R>(k, sm, tid);
}

template <int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) mbx_detect_kernel(const DetectParams p) {
    constexpr int T = NWARPS * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DSmem s;
    const bool nms = p.nms_iou >= 0.0f;
    const bool has_priors = p.priors != nullptr;   // NULL: the locations are absolute boxes already
    dcarve(&s, smem_raw, p.P, p.n2, p.k_max, nms, has_priors);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = p.P, N2 = p.n2, KM = p.k_max;
    const bool logits = (p.flags & MBX_FLAG_LOGITS) != 0;

    __shared__ HeadTab sh_heads[MBX_MAX_HEADS];
    const int nheads = p.nheads;
    if (nheads > 1) stage_heads(p, sh_heads);
    if (tid == 0) {
        mbar_init(s.bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0 && has_priors) {
        mbar_arrive_expect_tx(s.bar, static_cast<uint32_t>(sizeof(float4) * P));
        bulk_copy_g2s(s.priors, p.priors, static_cast<uint32_t>(sizeof(float4) * P), s.bar);
    }
    bool priors_ready = !has_priors;
    // Programmatic dependent launch (MBX_FLAG_PDL, see include/multibox_b200.h): this grid may start while
    // the preceding kernel of the stream (the previous detect step) is still running; everything it shares
    // with that kernel -- the output tensors -- is only written after griddepcontrol.wait.  Load, decode,
    // sort and NMS of step k+1 overlap the store phase and the completion of step k.
    const bool pdl = (p.flags & MBX_FLAG_PDL) != 0;
    bool dep_done = !pdl;
    if (pdl) asm volatile("griddepcontrol.launch_dependents;");
#ifdef MBX_PHASE_TIMING
    long long t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t_last = clock64();
#endif

    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        if (!priors_ready) {
            mbar_wait(s.bar, 0);
            priors_ready = true;
        }
        const size_t row0 = static_cast<size_t>(b) * P;
        float4 r = make_float4(0.f, 0.f, 1.f, 1.f);
        if (p.restrictions) r = reinterpret_cast<const float4 *>(p.restrictions)[b];
        if (tid == 0) {
            s.cnt[0] = 0;
            s.cnt[2] = 0;
        }
        for (int t = tid; t < kBins; t += T) s.hist[t] = 0;
        __syncthreads();
        // ---- decode + clip + restriction filter + sort key
        const float4 *gl = reinterpret_cast<const float4 *>(p.locations) + row0;
        int mine = 0;
        for (int j = tid; j < N2; j += T) {
            unsigned long long key = 0ull;
            if (j < P) {
                float4 l;
                float c;
                if (nheads > 1) {
                    int hh;
                    const size_t e = head_elem(sh_heads, nheads, j, b, hh);
                    l = ld_stream_f4(reinterpret_cast<const float4 *>(sh_heads[hh].loc) + e);
                    c = ld_stream_f(sh_heads[hh].conf + e);
                } else {
                    l = ld_stream_f4(gl + j);
                    c = ld_stream_f(p.confidences + row0 + j);
                }
                const float4 q = has_priors ? s.priors[j] : make_float4(0.f, 0.f, 0.f, 0.f);
                l.x = clip01(__fadd_rn(l.x, q.x));
                l.y = clip01(__fadd_rn(l.y, q.y));
                l.z = clip01(__fadd_rn(l.z, q.z));
                l.w = clip01(__fadd_rn(l.w, q.w));
                s.box[j] = l;
                if (logits) c = sigmoidf_(c);
                c = __fadd_rn(c, 0.0f);   // -0.0 -> +0.0 so equal values share one key
                // detect.py:92-99; a confidence of -inf marks a slot that holds no proposal (padding of
                // pooled candidate lists, multibox_b200/patches.py) -- no sigmoid output is ever -inf
                const bool drop = (l.x < r.x) || (l.y < r.y) || (l.z > r.z) || (l.w > r.w) || (c == -CUDART_INF_F);
                if (!drop) {
                    key = (static_cast<unsigned long long>(orderable(c)) << 32) | static_cast<unsigned>(j);
                    atomicAdd(&s.hist[conf_bin(c)], 1);
                    ++mine;
                }
            }
            s.keys[j] = key;
        }
        if (mine) atomicAdd(&s.cnt[0], mine);
        __syncthreads();
        MBX_DT(0);   // load / decode / keys
        // ---- pre-selection: only keys whose confidence bin reaches the bin of the keep-th largest
        // can be in the top max_to_keep; sort just those (typically ~keep of P) instead of all P.
        int keep = KM;
        if (p.max_to_keep) {
            keep = p.max_to_keep[b];
            keep = keep < 0 ? 0 : (keep > KM ? KM : keep);
        }
        const int n_valid = s.cnt[0];
        const int target = n_valid < keep ? n_valid : keep;
        if (warp == 0) {
            // lane l owns bins [hi-31, hi], hi = kBins-1-32*l (lane 0 = the largest confidences)
            const int hi = kBins - 1 - 32 * lane;
            int part = 0;
#pragma unroll 8
            for (int t = 0; t < 32; ++t) part += s.hist[hi - t];
            int cum = part;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, cum, o);
                if (lane >= o) cum += u;
            }
            const unsigned reach = __ballot_sync(0xffffffffu, cum >= target);
            int tb = 0;
            if (reach && target > 0) {
                const int l0 = __ffs(reach) - 1;
                if (lane == l0) {
                    int run = cum - part;
                    tb = hi;
                    for (int t = 0; t < 32; ++t) {
                        run += s.hist[hi - t];
                        tb = hi - t;
                        if (run >= target) break;
                    }
                }
                tb = __shfl_sync(0xffffffffu, tb, l0);
            }
            if (lane == 0) s.cnt[1] = (target > 0) ? tb : kBins;   // nothing to keep: select nothing
        }
        __syncthreads();
        const int tb = s.cnt[1];
        for (int j0 = 0; j0 < P; j0 += T) {
            const int j = j0 + tid;
            unsigned long long key = 0ull;
            bool take = false;
            if (j < P) {
                key = s.keys[j];
                take = key != 0ull && conf_bin(from_orderable(static_cast<uint32_t>(key >> 32))) >= tb;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, take);
            int base = 0;
            if (lane == 0 && bal) base = atomicAdd(&s.cnt[2], __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (take) s.ckeys[base + __popc(bal & ((1u << lane) - 1u))] = key;
        }
        __syncthreads();
        const int nc = s.cnt[2];
        int n2c = T;
        while (n2c < nc) n2c <<= 1;
        for (int t = nc + tid; t < n2c; t += T) s.ckeys[t] = 0ull;
        __syncthreads();
        // ---- descending sort of the candidates (registers + shuffles; shared memory across warps)
        if (n2c == T)
            sort_keys_in_smem<T, 1>(s.ckeys, tid);
        else if (n2c == 2 * T)
            sort_keys_in_smem<T, 2>(s.ckeys, tid);
        else if (n2c == 4 * T)
            sort_keys_in_smem<T, 4>(s.ckeys, tid);
        else if (n2c == 8 * T)
            sort_keys_in_smem<T, 8>(s.ckeys, tid);
        else {
            for (int k = 2; k <= n2c; k <<= 1) {      // generic fallback, all in shared memory
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int t = tid; t < (n2c >> 1); t += T) {
                        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                        const int ixj = i | j;
                        const unsigned long long a = s.ckeys[i], c = s.ckeys[ixj];
                        const bool desc = (i & k) == 0;
                        if ((a < c) == desc) {
                            s.ckeys[i] = c;
                            s.ckeys[ixj] = a;
                        }
                    }
                    __syncthreads();
                }
            }
        }
        MBX_DT(1);   // select + sort
        const int kk = target;
        int count = kk;
        // ---- greedy NMS over the kk sorted survivors (extension), chunk-serial bitmask formulation.
        // The survivors are cut into chunks of 32 in score order; lanes are boxes of a chunk.
        //   phase B  every chunk's OWN 32x32 triangle (does box i suppress a later box j of the same
        //            chunk?) is evaluated up front by all warps: one ballot per row -> diag[i];
        //   phase C  chunk by chunk: warp 0 resolves the chunk serially from diag and the bits earlier
        //            chunks left in remw[c] (32 shuffles + a 32-step register chain), then the warps
        //            share the ONLY pairwise work greedy NMS really needs -- the chunk's KEPT boxes
        //            against the boxes of the later chunks -- OR-ing their verdicts into remw[w]; two
        //            barriers per chunk.
        // Suppressed boxes are never used as suppressors, so about half of the upper triangle of
        // the pair matrix is never evaluated, and nothing is ever stored per pair.
        if (nms) {
            const int W = (kk + 31) >> 5;
            // a threshold outside [2^-20, 2^20] voids iou_fast's error analysis: decide every pair exactly
            const bool thr_ok = p.nms_iou >= 9.5367431640625e-07f && p.nms_iou <= 1048576.0f;
            for (int t = tid; t < (W << 5); t += T) {   // rows past kk: zero boxes (never kept, never dead)
                float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t < kk) bx = s.box[static_cast<unsigned>(s.ckeys[t] & 0xffffffffu)];
                s.sbox[t] = bx;
                s.sarea[t] = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
            }
            if (tid < 32) s.remw[tid] = 0u;
            __syncthreads();
            constexpr int RU = 8;   // rows per unit of phase B: RU independent IoU decisions per lane
            for (int t = warp; t < W * (32 / RU); t += NWARPS) {
                const int c = t / (32 / RU), ib = (t % (32 / RU)) * RU;
                const int i0 = (c << 5) + ib;
                if (i0 >= kk) continue;
                const int jj = (c << 5) + lane;
                const bool jvalid = jj < kk;
                const float4 bj = s.sbox[jj];
                const float aj = s.sarea[jj];
                const float4 *rbox = s.sbox + i0;
                const float *rarea = s.sarea + i0;
                unsigned bal[RU];
                bool any_near = !thr_ok && jvalid;
#pragma unroll
                for (int u = 0; u < RU; ++u) {
                    const bool act = jvalid && (lane > ib + u);   // j > i (and so i < kk, since j < kk)
                    bool nr;
                    const bool sp = iou_fast(rbox[u], rarea[u], bj, aj, p.nms_iou, nr);
                    bal[u] = __ballot_sync(0xffffffffu, sp && act);
                    any_near = any_near || (nr && act);
                }
                if (__any_sync(0xffffffffu, any_near)) {   // rare: a quotient within ulps of thr
#pragma unroll
                    for (int u = 0; u < RU; ++u) {
                        const bool act = jvalid && lane > ib + u;
                        bool nr;
                        bool sp = iou_fast(rbox[u], rarea[u], bj, aj, p.nms_iou, nr);
                        if (act && (nr || !thr_ok)) sp = iou_exact(rbox[u], rarea[u], bj, aj, p.nms_iou);
                        bal[u] = __ballot_sync(0xffffffffu, sp && act);
                    }
                }
                unsigned mine = 0u;
#pragma unroll
                for (int u = 0; u < RU; ++u) mine = (lane == u) ? bal[u] : mine;
                if (lane < RU) s.diag[i0 + lane] = mine;   // rows >= kk: all-zero ballots
            }
            __syncthreads();
            MBX_DT(2);   // diagonal triangles
            unsigned keptw = 0u;   // warp 0, lane c: kept mask of chunk c
            // serial resolve of chunk c by warp 0: 32 shuffles + a 32-step register chain; the kept rows go to
            // klist[buf] / cnt[4 + buf] (double buffered: the other warps may still be reading chunk c-1's list)
            auto resolve = [&](int c, int buf) {
                const unsigned d = s.diag[(c << 5) + lane];
                const int left = kk - (c << 5);
                const unsigned valid = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
                unsigned removed = s.remw[c] | ~valid;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const unsigned di = __shfl_sync(0xffffffffu, d, i);   // whom box i suppresses (bits > i)
                    // removed |= (bit i of removed clear) ? di : 0 -- as test-to-predicate + predicated OR:
                    // two dependent instructions per step on the 32-step chain instead of three
                    asm("{\n"
                        ".reg .pred p;\n"
                        ".reg .b32 t;\n"
                        "and.b32 t, %0, %2;\n"
                        "setp.eq.u32 p, t, 0;\n"
                        "@p or.b32 %0, %0, %1;\n"
                        "}\n"
                        : "+r"(removed)
                        : "r"(di), "r"(1u << i));
                }
                const unsigned kept = ~removed;
                if (lane == c) keptw = kept;
                const int nk = __popc(kept);
                // klist[l] = row of the l-th kept box of the chunk
                s.klist[buf * 32 + lane] = lane < nk ? (c << 5) + static_cast<int>(__fns(kept, 0, lane + 1)) : -1;
                if (lane == 0) s.cnt[4 + buf] = nk;
            };
            // The kept boxes of chunk c (list `buf`) against the target chunks w0 .. w0+nw-1, by the warps
            // wid = 0 .. nwp-1.  A lane owns NC boxes (one in each of NC target chunks: the row broadcast and
            // its bookkeeping are shared by NC decisions); the warps split the work by (group of NC chunks,
            // slice of the kept rows) and OR one ballot per chunk into remw[].
            auto cross = [&](auto nc_tag, int w0, int nw, int wid, int nwp, int buf) {
                constexpr int NC = decltype(nc_tag)::value;
                const int nkept = s.cnt[4 + buf];
                const int myrow = s.klist[buf * 32 + lane];
                const int CG = (nw + NC - 1) / NC;
                const int RS = CG >= nwp ? 1 : nwp / CG;
                for (int gg = wid; gg < CG * RS; gg += nwp) {
                    const int g = gg % CG, r = gg / CG;
                    float4 bj[NC];
                    float aj[NC];
                    bool jvalid[NC], dead[NC];
                    bool any_near = false;
#pragma unroll
                    for (int k = 0; k < NC; ++k) {
                        const int w = w0 + g * NC + k;
                        const int jj = (w << 5) + lane;
                        jvalid[k] = (g * NC + k) < nw && jj < kk;
                        bj[k] = s.sbox[jvalid[k] ? jj : 0];
                        aj[k] = s.sarea[jvalid[k] ? jj : 0];
                        dead[k] = false;
                        any_near = any_near || (!thr_ok && jvalid[k]);
                    }
                    for (int o = r; o < nkept; o += 2 * RS) {
                        const int i0 = __shfl_sync(0xffffffffu, myrow, o);
                        const bool on1 = o + RS < nkept;
                        const int i1 = on1 ? __shfl_sync(0xffffffffu, myrow, (o + RS) & 31) : i0;
                        const float4 b0 = s.sbox[i0], b1 = s.sbox[i1];
                        const float a0 = s.sarea[i0], a1 = s.sarea[i1];
#pragma unroll
                        for (int k = 0; k < NC; ++k) {
                            bool n0, n1;
                            const bool s0 = iou_fast(b0, a0, bj[k], aj[k], p.nms_iou, n0);
                            const bool s1 = iou_fast(b1, a1, bj[k], aj[k], p.nms_iou, n1);
                            dead[k] = dead[k] || s0 || s1;      // (i1 == i0 when the second row is off)
                            any_near = any_near || ((n0 || n1) && jvalid[k]);
                        }
                    }
                    if (__any_sync(0xffffffffu, any_near)) {   // rare: redo my rows with the IEEE division
                        for (int k = 0; k < NC; ++k) dead[k] = false;
                        for (int o = r; o < nkept; o += RS) {
                            const int i = __shfl_sync(0xffffffffu, myrow, o);
#pragma unroll
                            for (int k = 0; k < NC; ++k) {
                                bool nr;
                                bool sp = iou_fast(s.sbox[i], s.sarea[i], bj[k], aj[k], p.nms_iou, nr);
                                if (nr || !thr_ok) sp = iou_exact(s.sbox[i], s.sarea[i], bj[k], aj[k], p.nms_iou);
                                dead[k] = dead[k] || sp;
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < NC; ++k) {
                        const unsigned m = __ballot_sync(0xffffffffu, dead[k] && jvalid[k]);
                        if (lane == 0 && m) atomicOr(&s.remw[w0 + g * NC + k], m);
                    }
                }
            };
            // boxes per lane: the split that leaves the fewest empty (chunk, lane) slots
            auto cross_any = [&](int w0, int nw, int wid, int nwp, int buf) {
                if (nw >= 7 || nw == 4)
                    cross(std::integral_constant<int, 4>{}, w0, nw, wid, nwp, buf);
                else if (nw >= 3)       // 6 = 3+3, 5 = 3+2, 3
                    cross(std::integral_constant<int, 3>{}, w0, nw, wid, nwp, buf);
                else if (nw == 2)
                    cross(std::integral_constant<int, 2>{}, w0, nw, wid, nwp, buf);
                else if (nw == 1)
                    cross(std::integral_constant<int, 1>{}, w0, nw, wid, nwp, buf);
            };
            // Schedule.  Chunk c+1 can be resolved as soon as the kept boxes of chunks <= c have been tested
            // against IT -- not against the chunks after it.  So per chunk c: (A) every warp tests chunk c's
            // kept boxes against chunk c+1 only; barrier; (B) warp 0 resolves chunk c+1 (the serial 32-step
            // chain) WHILE the other warps test chunk c's kept boxes against chunks c+2 ..; barrier.
            if (warp == 0) resolve(0, 0);
            MBX_DT(5);   // (timing builds) serial resolve
            __syncthreads();
            MBX_DT(6);   // (timing builds) wait for the resolve
            for (int c = 0; c + 1 < W; ++c) {
                const int buf = c & 1;
                cross_any(c + 1, 1, warp, NWARPS, buf);                                   // (A)
                MBX_DT(7);   // (timing builds) cross-chunk suppression
                __syncthreads();   // remw[c+1] complete
                MBX_DT(6);
                if (warp == 0) {                                                          // (B)
                    resolve(c + 1, buf ^ 1);
                    MBX_DT(5);
                } else {
                    cross_any(c + 2, W - 2 - c, warp - 1, NWARPS - 1, buf);
                    MBX_DT(7);
                }
                __syncthreads();   // list of chunk c+1 and remw[c+2..] (from chunk c) complete
                MBX_DT(6);
            }
            if (warp == 0) {
                // exclusive prefix of kept counts per word
                const int c = __popc(keptw);
                int inc = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += t;
                }
                s.keepw[lane] = keptw;
                s.keeppre[lane] = inc - c;
                if (lane == 31) s.keeppre[32] = inc;
            }
            __syncthreads();
            count = s.keeppre[32];
            MBX_DT(3);   // chunk-serial resolve + cross-chunk suppression
        }
        // ---- store (convert_proposals in float64)
        if (!dep_done) {   // first global write of this CTA: the preceding grid must be complete
            asm volatile("griddepcontrol.wait;" ::: "memory");
            dep_done = true;
        }
        double sx = 1.0, sy = 1.0, ox = 0.0, oy = 0.0;
        int flip = 0;
        if (p.image_dims) {
            const double ih = static_cast<double>(p.image_dims[2 * b]), iw = static_cast<double>(p.image_dims[2 * b + 1]);
            sx = __ddiv_rn(static_cast<double>(p.patch_dims[2 * b + 1]), iw);
            sy = __ddiv_rn(static_cast<double>(p.patch_dims[2 * b]), ih);
            ox = __ddiv_rn(static_cast<double>(p.offsets[2 * b + 1]), iw);
            oy = __ddiv_rn(static_cast<double>(p.offsets[2 * b]), ih);
            flip = p.is_flipped ? p.is_flipped[b] : 0;
        }
        const size_t out0 = static_cast<size_t>(b) * KM;
        if (!nms) {
            for (int t = tid; t < KM; t += T) {
                float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
                float sc = 0.f;
                int idx = -1;
                if (t < kk) {
                    const unsigned long long key = s.ckeys[t];
                    idx = static_cast<int>(key & 0xffffffffu);
                    sc = from_orderable(static_cast<uint32_t>(key >> 32));
                    bx = s.box[idx];
                }
                if (p.out_patch_boxes) reinterpret_cast<float4 *>(p.out_patch_boxes)[out0 + t] = bx;
                if (p.out_scores) p.out_scores[out0 + t] = sc;
                if (p.out_idx) p.out_idx[out0 + t] = idx;
                if (p.out_boxes) {
                    double x1 = 0., y1 = 0., x2 = 0., y2 = 0.;
                    if (t < kk) {
                        x1 = __dadd_rn(__dmul_rn(static_cast<double>(bx.x), sx), ox);
                        y1 = __dadd_rn(__dmul_rn(static_cast<double>(bx.y), sy), oy);
                        x2 = __dadd_rn(__dmul_rn(static_cast<double>(bx.z), sx), ox);
                        y2 = __dadd_rn(__dmul_rn(static_cast<double>(bx.w), sy), oy);
                        if (flip) {
                            const double t1 = __dsub_rn(1.0, x2), t2 = __dsub_rn(1.0, x1);
                            x1 = t1;
                            x2 = t2;
                        }
                    }
                    double2 *ob = reinterpret_cast<double2 *>(p.out_boxes) + 2 * (out0 + t);
                    ob[0] = make_double2(x1, y1);
                    ob[1] = make_double2(x2, y2);
                }
            }
        } else {
            // kept entries scatter to their rank; the tail [count, KM) is zero-filled
            for (int t = tid; t < KM; t += T) {
                const bool in_range = t < kk;
                bool kept = false;
                int pos = 0;
                if (in_range) {
                    const unsigned wbits = s.keepw[t >> 5];
                    kept = (wbits >> (t & 31)) & 1u;
                    pos = s.keeppre[t >> 5] + __popc(wbits & ((1u << (t & 31)) - 1u));
                }
                if (kept) {
                    const unsigned long long key = s.ckeys[t];
                    const int idx = static_cast<int>(key & 0xffffffffu);
                    const float sc = from_orderable(static_cast<uint32_t>(key >> 32));
                    const float4 bx = s.sbox[t];
                    if (p.out_patch_boxes) reinterpret_cast<float4 *>(p.out_patch_boxes)[out0 + pos] = bx;
                    if (p.out_scores) p.out_scores[out0 + pos] = sc;
                    if (p.out_idx) p.out_idx[out0 + pos] = idx;
                    if (p.out_boxes) {
                        double x1 = __dadd_rn(__dmul_rn(static_cast<double>(bx.x), sx), ox);
                        double y1 = __dadd_rn(__dmul_rn(static_cast<double>(bx.y), sy), oy);
                        double x2 = __dadd_rn(__dmul_rn(static_cast<double>(bx.z), sx), ox);
                        double y2 = __dadd_rn(__dmul_rn(static_cast<double>(bx.w), sy), oy);
                        if (flip) {
                            const double t1 = __dsub_rn(1.0, x2), t2 = __dsub_rn(1.0, x1);
                            x1 = t1;
                            x2 = t2;
                        }
                        double2 *ob = reinterpret_cast<double2 *>(p.out_boxes) + 2 * (out0 + pos);
                        ob[0] = make_double2(x1, y1);
                        ob[1] = make_double2(x2, y2);
                    }
                }
                if (t >= count) {
                    if (p.out_patch_boxes)
                        reinterpret_cast<float4 *>(p.out_patch_boxes)[out0 + t] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.out_scores) p.out_scores[out0 + t] = 0.f;
                    if (p.out_idx) p.out_idx[out0 + t] = -1;
                    if (p.out_boxes) {
                        double2 *ob = reinterpret_cast<double2 *>(p.out_boxes) + 2 * (out0 + t);
                        ob[0] = make_double2(0., 0.);
                        ob[1] = make_double2(0., 0.);
                    }
                }
            }
        }
        if (tid == 0 && p.out_count) p.out_count[b] = count;
        __syncthreads();   // shared state is reused by the next image
        MBX_DT(4);   // store
    }
    if (!dep_done) asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef MBX_PHASE_TIMING
    if (lane == 0 && p.out_scores) {
        long long *dbg = reinterpret_cast<long long *>(p.out_scores) + (static_cast<size_t>(blockIdx.x) * NWARPS + warp) * 8;
        for (int k = 0; k < 8; ++k) dbg[k] = t_acc[k];
    }
#endif
}

template <int NWARPS>
static int launch_detect(const DetectParams &p, size_t smem, cudaStream_t st) {
    auto kern = mbx_detect_kernel<NWARPS>;
    static thread_local size_t configured = 0;
    static thread_local int occ = 0;
    static thread_local int cached_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {   // function attributes and occupancy are per device
        configured = 0;
        occ = 0;
        cached_dev = dev;
    }
    if (smem > configured || occ == 0) {
        if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    static_cast<int>(smem)),
                               "cudaFuncSetAttribute(detect)"))
            return e;
        configured = smem;
        occ = 0;
    }
    static thread_local size_t occ_smem = 0;
    if (occ == 0 || occ_smem != smem) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NWARPS * 32, smem);
        if (occ < 1) occ = 1;
        occ_smem = smem;
    }
    int grid = sm_count() * occ;
    if (grid > p.B) grid = p.B;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NWARPS * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (p.flags & MBX_FLAG_PDL) ? 1 : 0;
    return check_cuda(cudaLaunchKernelEx(&cfg, kern, p), "launch mbx_detect_kernel");
}


// ---- batched filter_proposals (detect.py:74-104): ordered compaction, one CTA per image
__global__ void __launch_bounds__(256) mbx_filter_kernel(const float *bboxes, const float *conf,
                                                         const float *restrictions, int B, int P,
                                                         float *out_b, float *out_c, int32_t *out_idx,
                                                         int32_t *out_count) {
    __shared__ int warp_cnt[8];
    __shared__ int base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        float4 r = make_float4(0.1f, 0.1f, 0.9f, 0.9f);   // the reference's default
        if (restrictions) r = reinterpret_cast<const float4 *>(restrictions)[b];
        const size_t row0 = static_cast<size_t>(b) * P;
        if (tid == 0) base = 0;
        __syncthreads();
        for (int j0 = 0; j0 < P; j0 += 256) {
            const int j = j0 + tid;
            bool keep = false;
            float4 l = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < P) {
                l = ld_stream_f4(reinterpret_cast<const float4 *>(bboxes) + row0 + j);
                keep = !((l.x < r.x) || (l.y < r.y) || (l.z > r.z) || (l.w > r.w));
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) warp_cnt[warp] = __popc(bal);
            __syncthreads();
            int off = base;
            for (int w = 0; w < warp; ++w) off += warp_cnt[w];
            int tot = 0;
            for (int w = 0; w < 8; ++w) tot += warp_cnt[w];
            if (keep) {
                const int pos = off + __popc(bal & ((1u << lane) - 1u));
                reinterpret_cast<float4 *>(out_b)[row0 + pos] = l;
                out_c[row0 + pos] = conf[row0 + j];
                if (out_idx) out_idx[row0 + pos] = j;
            }
            __syncthreads();
            if (tid == 0) base += tot;
            __syncthreads();
        }
        if (tid == 0) out_count[b] = base;
        __syncthreads();
    }
}

// ---- batched convert_proposals (detect.py:106-131), float64
__global__ void __launch_bounds__(256) mbx_convert_kernel(const float *bboxes, const int32_t *offsets,
                                                          const int32_t *patch_dims, const int32_t *image_dims,
                                                          const int32_t *is_flipped, const int32_t *counts, int B,
                                                          int K, double *out) {
    const long long total = static_cast<long long>(B) * K;
    for (long long e = blockIdx.x * 256ll + threadIdx.x; e < total; e += 256ll * gridDim.x) {
        const int b = static_cast<int>(e / K), t = static_cast<int>(e - static_cast<long long>(b) * K);
        double x1 = 0., y1 = 0., x2 = 0., y2 = 0.;
        if (!counts || t < counts[b]) {
            const float4 bx = reinterpret_cast<const float4 *>(bboxes)[e];
            const double ih = static_cast<double>(image_dims[2 * b]), iw = static_cast<double>(image_dims[2 * b + 1]);
            const double sx = __ddiv_rn(static_cast<double>(patch_dims[2 * b + 1]), iw);
            const double sy = __ddiv_rn(static_cast<double>(patch_dims[2 * b]), ih);
            const double ox = __ddiv_rn(static_cast<double>(offsets[2 * b + 1]), iw);
            const double oy = __ddiv_rn(static_cast<double>(offsets[2 * b]), ih);
            x1 = __dadd_rn(__dmul_rn(static_cast<double>(bx.x), sx), ox);
            y1 = __dadd_rn(__dmul_rn(static_cast<double>(bx.y), sy), oy);
            x2 = __dadd_rn(__dmul_rn(static_cast<double>(bx.z), sx), ox);
            y2 = __dadd_rn(__dmul_rn(static_cast<double>(bx.w), sy), oy);
            if (is_flipped && is_flipped[b]) {
                const double t1 = __dsub_rn(1.0, x2), t2 = __dsub_rn(1.0, x1);
                x1 = t1;
                x2 = t2;
            }
        }
        double2 *ob = reinterpret_cast<double2 *>(out) + 2 * e;
        ob[0] = make_double2(x1, y1);
        ob[1] = make_double2(x2, y2);
    }
}

}  // namespace mbx

using namespace mbx;

extern "C" int mbx_filter_proposals(const float *bboxes, const float *confidences, const float *restrictions,
                                    int B, int P, float *out_bboxes, float *out_confidences, int32_t *out_idx,
                                    int32_t *out_count, void *stream) {
    if (B < 0 || P <= 0 || !bboxes || !confidences || !out_bboxes || !out_confidences || !out_count) {
        set_error("mbx_filter_proposals: bad arguments");
        return MBX_E_ARG;
    }
    if (B == 0) return 0;
    int grid = B < sm_count() * 8 ? B : sm_count() * 8;
    mbx_filter_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(bboxes, confidences, restrictions, B, P,
                                                                           out_bboxes, out_confidences, out_idx,
                                                                           out_count);
    return check_cuda(cudaGetLastError(), "launch mbx_filter_kernel");
}

extern "C" int mbx_convert_proposals(const float *bboxes, const int32_t *offsets, const int32_t *patch_dims,
                                     const int32_t *image_dims, const int32_t *is_flipped, const int32_t *counts,
                                     int B, int K, double *out_boxes, void *stream) {
    if (B < 0 || K < 0 || !bboxes || !offsets || !patch_dims || !image_dims || !out_boxes) {
        set_error("mbx_convert_proposals: bad arguments");
        return MBX_E_ARG;
    }
    if (B == 0 || K == 0) return 0;
    long long total = static_cast<long long>(B) * K;
    long long blocks = (total + 255) / 256;
    int grid = static_cast<int>(blocks < sm_count() * 16ll ? blocks : sm_count() * 16ll);
    mbx_convert_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(bboxes, offsets, patch_dims, image_dims,
                                                                            is_flipped, counts, B, K, out_boxes);
    return check_cuda(cudaGetLastError(), "launch mbx_convert_kernel");
}

extern "C" size_t mbx_detect_workspace_bytes(int B, int P, int k_max) {
    (void)B;
    (void)P;
    (void)k_max;
    return 256;   // the detect path needs no global scratch; kept in the ABI for symmetry
}

namespace mbx {
static int detect_impl(const mbx_heads *heads, const float *locations, const float *confidences, const float *priors,
                       const float *restrictions, const int32_t *max_to_keep, const int32_t *offsets,
                       const int32_t *patch_dims, const int32_t *image_dims, const int32_t *is_flipped, int B, int P,
                       int k_max, float nms_iou, unsigned flags, double *out_boxes, float *out_patch_boxes,
                       float *out_scores, int32_t *out_prior_idx, int32_t *out_count, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (B < 0 || P <= 0 || k_max <= 0) {
        set_error("mbx_detect: bad sizes B=%d P=%d k_max=%d", B, P, k_max);
        return MBX_E_ARG;
    }
    if (B == 0) return 0;
    if (!heads && (!locations || !confidences)) {
        set_error("mbx_detect: null input pointer");
        return MBX_E_ARG;
    }
    if (heads) {
        int tot = 0;
        bool bad = heads->num_heads < 1 || heads->num_heads > MBX_MAX_HEADS;
        for (int k = 0; !bad && k < heads->num_heads; ++k) {
            bad = heads->head_priors[k] <= 0 || !heads->locations[k] || !heads->confidences[k] ||
                  (reinterpret_cast<uintptr_t>(heads->locations[k]) & 15u);
            tot += heads->head_priors[k];
        }
        if (bad || tot != P) {
            set_error("mbx_detect_heads: bad head table (num_heads=%d, sum of head_priors=%d, P=%d)",
                      heads->num_heads, tot, P);
            return MBX_E_ARG;
        }
    }
    if ((image_dims != nullptr) != (patch_dims != nullptr) || (image_dims != nullptr) != (offsets != nullptr)) {
        set_error("mbx_detect: offsets, patch_dims and image_dims must be given together");
        return MBX_E_ARG;
    }
    auto mis16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15u) != 0; };
    if (mis16(locations) || mis16(priors) || mis16(restrictions) || mis16(out_boxes) || mis16(out_patch_boxes)) {
        set_error("mbx_detect: pointers must be 16-byte aligned");
        return MBX_E_ARG;
    }
    if (k_max > 1024) {
        set_error("mbx_detect: k_max=%d > 1024", k_max);
        return MBX_E_TOO_LARGE;
    }
    DetectParams p;
    p.locations = locations;
    p.confidences = confidences;
    p.nheads = 1;
    for (int k = 0; k < MBX_MAX_HEADS; ++k) p.heads[k] = HeadTab{nullptr, nullptr, nullptr, nullptr, 0, 0};
    if (heads) {
        int off = 0;
        p.nheads = heads->num_heads;
        for (int k = 0; k < heads->num_heads; ++k) {
            p.heads[k] = HeadTab{heads->locations[k], heads->confidences[k], nullptr, nullptr, heads->head_priors[k], off};
            off += heads->head_priors[k];
        }
        if (p.nheads == 1) {
            p.locations = heads->locations[0];
            p.confidences = heads->confidences[0];
        }
    }
    p.priors = priors;
    p.restrictions = restrictions;
    p.max_to_keep = max_to_keep;
    p.offsets = offsets;
    p.patch_dims = patch_dims;
    p.image_dims = image_dims;
    p.is_flipped = is_flipped;
    p.B = B;
    p.P = P;
    p.k_max = k_max;
    int nwarps = static_cast<int>((flags >> MBX_FLAG_WARPS_SHIFT) & 0xffu);
    if (nwarps == 0) nwarps = 8;
    int n2 = nwarps * 32;            // at least one key per thread (register/shuffle sort)
    while (n2 < P) n2 <<= 1;
    p.n2 = n2;
    p.nms_iou = nms_iou;
    p.flags = flags;
    p.out_boxes = out_boxes;
    p.out_patch_boxes = out_patch_boxes;
    p.out_scores = out_scores;
    p.out_idx = out_prior_idx;
    p.out_count = out_count;
    const size_t smem = dcarve(nullptr, nullptr, P, n2, k_max, nms_iou >= 0.0f, priors != nullptr);
    if (smem > static_cast<size_t>(max_smem_optin())) {
        set_error("mbx_detect: P=%d k_max=%d needs %zu bytes of shared memory per CTA (max %d)", P, k_max, smem,
                  max_smem_optin());
        return MBX_E_TOO_LARGE;
    }
    switch (nwarps) {
        case 4: return launch_detect<4>(p, smem, st);
        case 8: return launch_detect<8>(p, smem, st);
        case 16: return launch_detect<16>(p, smem, st);
        default:
            set_error("mbx_detect: forced warps must be 4, 8 or 16");
            return MBX_E_ARG;
    }
}
}  // namespace mbx

extern "C" int mbx_detect(const float *locations, const float *confidences, const float *priors,
                          const float *restrictions, const int32_t *max_to_keep, const int32_t *offsets,
                          const int32_t *patch_dims, const int32_t *image_dims, const int32_t *is_flipped, int B,
                          int P, int k_max, float nms_iou, unsigned flags, double *out_boxes,
                          float *out_patch_boxes, float *out_scores, int32_t *out_prior_idx, int32_t *out_count,
                          void *workspace, size_t workspace_bytes, void *stream) {
    (void)workspace;
    (void)workspace_bytes;
    return detect_impl(nullptr, locations, confidences, priors, restrictions, max_to_keep, offsets, patch_dims,
                       image_dims, is_flipped, B, P, k_max, nms_iou, flags, out_boxes, out_patch_boxes, out_scores,
                       out_prior_idx, out_count, stream);
}

extern "C" int mbx_detect_heads(const mbx_heads *heads, const float *priors, const float *restrictions,
                                const int32_t *max_to_keep, const int32_t *offsets, const int32_t *patch_dims,
                                const int32_t *image_dims, const int32_t *is_flipped, int B, int P, int k_max,
                                float nms_iou, unsigned flags, double *out_boxes, float *out_patch_boxes,
                                float *out_scores, int32_t *out_prior_idx, int32_t *out_count, void *workspace,
                                size_t workspace_bytes, void *stream) {
    (void)workspace;
    (void)workspace_bytes;
    if (!heads) {
        set_error("mbx_detect_heads: null heads");
        return MBX_E_ARG;
    }
    return detect_impl(heads, nullptr, nullptr, priors, restrictions, max_to_keep, offsets, patch_dims, image_dims,
                       is_flipped, B, P, k_max, nms_iou, flags, out_boxes, out_patch_boxes, out_scores, out_prior_idx,
                       out_count, stream);
}
