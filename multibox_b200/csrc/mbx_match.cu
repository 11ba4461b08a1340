// multibox_b200 -- GT->prior optimal matching fused with the multibox loss
// forward/backward, one CTA per image (persistent grid), sm_100a.
//
// What it replaces (reference = gvanhorn38/multibox):
//   loss.py:21-25   log terms                      -> nplogf() at load time
//   loss.py:33-35   cost matrix C[P, n_b] (fp64)   -> never materialised: C(p, j) is
//                                                     recomputed in registers with the
//                                                     reference's fp32 operation order
//   loss.py:40      scipy linear_sum_assignment    -> shortest-augmenting-path solver
//                                                     (Crouse 2016, the algorithm scipy >= 1.4
//                                                     ships) on the transposed problem
//                                                     (rows = GT, columns = priors), same
//                                                     scan order and tie rule
//   loss.py:44-53   mask + stacked GT              -> epilogue
//   loss.py:67-74   prior add, epsilon add         -> at load time
//   loss.py:88-101  partition + the two losses     -> epilogue (fp64 accumulation)
//   TF autodiff     gradients                      -> epilogue
//   model.py:322    sigmoid (MBX_FLAG_LOGITS)      -> at load time
//
// Data layout in shared memory (per CTA, one image at a time):
//   priors   float4[P]  staged ONCE per CTA with a TMA bulk copy (cp.async.bulk)
//   loc      float4[P]  absolute predicted boxes (offset + prior)
//   spc      double[P]  shortest path cost of each column in the current augmentation
//   v        double[P]  column duals
//   conf/lc/l1 float[P] confidence (+eps), log(c), log(clamp(1-c))
//   path,row4col int16[P]; sc uint8[P] (column already scanned in this augmentation)
//   gt float4[M], u double[M], col4row int[M], removal log int[M] x2
// Every thread owns the columns j = tid, tid+T, ... so all per-column traffic is
// conflict-free and needs no barrier; one __syncthreads per Dijkstra step
// (the block-wide arg-min) and three per augmentation.
#include <unordered_map>
#include <new>

#include "mbx_match.cuh"

namespace mbx {

constexpr int kTieBit = 1 << 30;
constexpr int kNoCol = kTieBit - 1;

struct Cand {
    double v;
    int j;   // column | kTieBit
};

__device__ __forceinline__ Cand cand_min(Cand a, Cand b) {
    if (b.v < a.v) return b;
    if (a.v < b.v) return a;
    int ja = a.j & ~kTieBit, jb = b.j & ~kTieBit;
    Cand r;
    r.v = a.v;
    r.j = (ja < jb ? ja : jb) | kTieBit;
    return r;
}

__device__ __forceinline__ Cand warp_cand_min(Cand c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Cand t;
        t.v = __shfl_xor_sync(0xffffffffu, c.v, o);
        t.j = __shfl_xor_sync(0xffffffffu, c.j, o);
        c = cand_min(c, t);
    }
    return c;
}

struct Smem {
    float4 *priors, *loc, *gt;
    double *spc, *v, *u, *red;
    float *conf, *lc, *l1;
    int *col4row, *rm_col, *rm_idx, *pj, *ri;
    double *pv;
    unsigned long long *pk;
    short *path, *row4col;
    unsigned char *sc;
    uint64_t *bar;
};

// Shared-memory carve-up; identical on host (size query) and device.
__host__ __device__ inline size_t carve(Smem *s, unsigned char *base, int P, int M, int nwarps, bool has_priors) {
    size_t o = 0;
    auto take = [&](size_t bytes, size_t al) {
        o = align_up(o, al);
        size_t r = o;
        o += bytes;
        return r;
    };
    size_t o_pri = take(has_priors ? sizeof(float4) * P : 0, 16);
    size_t o_loc = take(sizeof(float4) * P, 16);
    size_t o_gt = take(sizeof(float4) * (M > 0 ? M : 1), 16);
    size_t o_spc = take(sizeof(double) * P, 8);
    size_t o_v = take(sizeof(double) * P, 8);
    size_t o_u = take(sizeof(double) * (M > 0 ? M : 1), 8);
    size_t o_red = take(sizeof(double) * 2 * nwarps, 8);
    size_t o_pv = take(sizeof(double) * 2 * nwarps, 8);
    size_t o_pk = take(sizeof(unsigned long long) * nwarps, 8);
    size_t o_bar = take(sizeof(uint64_t), 8);
    size_t o_conf = take(sizeof(float) * P, 4);
    size_t o_lc = take(sizeof(float) * P, 4);
    size_t o_l1 = take(sizeof(float) * P, 4);
    size_t o_c4r = take(sizeof(int) * (M > 0 ? M : 1), 4);
    size_t o_rmc = take(sizeof(int) * (M + 1), 4);
    size_t o_rmi = take(sizeof(int) * (M + 1), 4);
    size_t o_pj = take(sizeof(int) * 2 * nwarps, 4);
    size_t o_ri = take(sizeof(int) * nwarps, 4);
    size_t o_path = take(sizeof(short) * P, 2);
    size_t o_r4c = take(sizeof(short) * P, 2);
    size_t o_sc = take(P, 1);
    if (s) {
        s->priors = reinterpret_cast<float4 *>(base + o_pri);
        s->loc = reinterpret_cast<float4 *>(base + o_loc);
        s->gt = reinterpret_cast<float4 *>(base + o_gt);
        s->spc = reinterpret_cast<double *>(base + o_spc);
        s->v = reinterpret_cast<double *>(base + o_v);
        s->u = reinterpret_cast<double *>(base + o_u);
        s->red = reinterpret_cast<double *>(base + o_red);
        s->pv = reinterpret_cast<double *>(base + o_pv);
        s->pk = reinterpret_cast<unsigned long long *>(base + o_pk);
        s->bar = reinterpret_cast<uint64_t *>(base + o_bar);
        s->conf = reinterpret_cast<float *>(base + o_conf);
        s->lc = reinterpret_cast<float *>(base + o_lc);
        s->l1 = reinterpret_cast<float *>(base + o_l1);
        s->col4row = reinterpret_cast<int *>(base + o_c4r);
        s->rm_col = reinterpret_cast<int *>(base + o_rmc);
        s->rm_idx = reinterpret_cast<int *>(base + o_rmi);
        s->pj = reinterpret_cast<int *>(base + o_pj);
        s->ri = reinterpret_cast<int *>(base + o_ri);
        s->path = reinterpret_cast<short *>(base + o_path);
        s->row4col = reinterpret_cast<short *>(base + o_r4c);
        s->sc = base + o_sc;
    }
    return align_up(o, 16);
}

// exclusive scan of clamp(num_gt, 0, M) -> offsets[B+1]; one CTA of 1024 threads.
__global__ void __launch_bounds__(1024) mbx_scan_num_gt_kernel(const int32_t *num_gt, const int32_t *gt_row, int B,
                                                               int M, int32_t *offsets, int32_t *n_stacked) {
    __shared__ int warp_excl[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < B; base += 1024) {
        const int i = base + tid;
        int x = 0;
        if (i < B) {
            x = image_num_gt(num_gt, gt_row, i);
            x = x < 0 ? 0 : (x > M ? M : x);
        }
        int inc = x;   // inclusive scan inside the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_excl[w] = inc;
        __syncthreads();
        if (w == 0) {
            const int tot = warp_excl[lane];
            int ti = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int q = __shfl_up_sync(0xffffffffu, ti, o);
                if (lane >= o) ti += q;
            }
            warp_excl[lane] = ti - tot;
        }
        __syncthreads();
        const int excl = carry + warp_excl[w] + inc - x;
        if (i < B) offsets[i] = excl;
        __syncthreads();
        if (tid == 1023) carry = excl + x;
        __syncthreads();
    }
    if (tid == 0) {
        offsets[B] = carry;
        if (n_stacked) *n_stacked = carry;
    }
}

template <int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) mbx_match_loss_kernel(const MatchParams p) {
    constexpr int T = NWARPS * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem s;
    const bool boundary = (p.flags & MBX_FLAG_BOUNDARY) != 0;
    const bool logits = (p.flags & MBX_FLAG_LOGITS) != 0;
    const bool has_priors = !boundary;
    carve(&s, smem_raw, p.P, p.M, NWARPS, has_priors);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = p.P, M = p.M;
    const float half_alpha = __fdiv_rn(p.alpha, 2.0f);   // (alpha / 2.) in fp32, loss.py:35
    const double INF = CUDART_INF;
    unsigned status = 0;

    // ---- priors: one TMA bulk copy per CTA, reused for every image this CTA solves
    if (has_priors) {
        if (tid == 0) {
            mbar_init(s.bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (tid == 0) {
            mbar_arrive_expect_tx(s.bar, static_cast<uint32_t>(sizeof(float4) * P));
            bulk_copy_g2s(s.priors, p.priors, static_cast<uint32_t>(sizeof(float4) * P), s.bar);
        }
    }
    bool priors_ready = !has_priors;
    int pbuf = 0;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
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
        // ---- load + elementwise prologue (loss.py:67-74, 21-25; model.py:322)
        const float4 *gl = reinterpret_cast<const float4 *>(p.locations) + row0;
        for (int j = tid; j < P; j += T) {
            float4 l = ld_stream_f4(gl + j);
            if (has_priors) {
                float4 q = s.priors[j];
                l.x = __fadd_rn(l.x, q.x);
                l.y = __fadd_rn(l.y, q.y);
                l.z = __fadd_rn(l.z, q.z);
                l.w = __fadd_rn(l.w, q.w);
            }
            s.loc[j] = l;
            float c = ld_stream_f(p.confidences + row0 + j);
            if (logits) {
                c = sigmoidf_(c);
                if (p.conf_out) p.conf_out[row0 + j] = c;
            }
            s.conf[j] = c;                                  // pre-epsilon value (sigmoid output)
            if (!boundary) c = __fadd_rn(c, kEps32);        // loss.py:74
            s.lc[j] = nplogf(c);
            float v = __fsub_rn(1.0f, c);
            if (v > 1.0f) v = 1.0f;
            if (v <= 0.0f) v = kEps32;
            s.l1[j] = nplogf(v);
            s.v[j] = 0.0;
            s.row4col[j] = -1;
            s.sc[j] = 0;
        }
        for (int i = tid; i < n; i += T) {
            s.gt[i] = gg[i];
            s.u[i] = 0.0;
            s.col4row[i] = -1;
        }
        __syncthreads();

        // ---- one shortest augmenting path per GT row
        bool failed = false;
        for (int cur = 0; cur < n && !failed; ++cur) {
            int i = cur, R = 0;
            double min_val = 0.0;
            for (;;) {
                const float4 g = s.gt[i];
                const double ui = s.u[i];
                const bool first = (R == 0);
                Cand best;
                best.v = INF;
                best.j = kNoCol;
                for (int j = tid; j < P; j += T) {
                    if (s.sc[j]) continue;
                    float c32 = cost32(s.loc[j], g, half_alpha, s.lc[j], s.l1[j]);
                    if (c32 != c32 || c32 == -CUDART_INF_F) status |= MBX_STATUS_INVALID_COST;
                    double r = __dsub_rn(__dsub_rn(__dadd_rn(min_val, static_cast<double>(c32)), ui), s.v[j]);
                    double sp = first ? INF : s.spc[j];
                    bool upd = r < sp;
                    if (upd) {
                        sp = r;
                        s.path[j] = static_cast<short>(i);
                    }
                    if (upd || first) s.spc[j] = sp;
                    if (sp < best.v) {
                        best.v = sp;
                        best.j = j;
                    } else if (sp == best.v) {
                        best.j |= kTieBit;
                    }
                }
                // block-wide arg-min (value, lowest column, tie flag); one barrier
                best = warp_cand_min(best);
                if (NWARPS > 1) {
                    if (lane == 0) {
                        s.pv[pbuf * NWARPS + warp] = best.v;
                        s.pj[pbuf * NWARPS + warp] = best.j;
                    }
                    __syncthreads();
                    best.v = s.pv[pbuf * NWARPS];
                    best.j = s.pj[pbuf * NWARPS];
#pragma unroll
                    for (int w = 1; w < NWARPS; ++w) {
                        Cand t;
                        t.v = s.pv[pbuf * NWARPS + w];
                        t.j = s.pj[pbuf * NWARPS + w];
                        best = cand_min(best, t);
                    }
                    pbuf ^= 1;
                }
                min_val = best.v;
                if (!(min_val < INF)) {   // infeasible (scipy raises ValueError)
                    status |= MBX_STATUS_INFEASIBLE;
                    failed = true;
                    break;
                }
                int jstar = best.j & ~kTieBit;
                if (best.j & kTieBit) {
                    // scipy's tie rule (rare path): among the columns at the minimum, the LAST
                    // unassigned one in `remaining` order wins, else the FIRST assigned one.
                    unsigned long long key = ~0ull;
                    for (int j = tid; j < P; j += T) {
                        if (s.sc[j] || !(s.spc[j] == min_val)) continue;
                        int pos = replay_pos(j, R, P, s.rm_idx);
                        unsigned k2 = (s.row4col[j] < 0) ? static_cast<unsigned>(P - 1 - pos)
                                                         : static_cast<unsigned>(P + pos);
                        unsigned long long k = (static_cast<unsigned long long>(k2) << 32) | static_cast<unsigned>(j);
                        key = k < key ? k : key;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        unsigned long long t = __shfl_xor_sync(0xffffffffu, key, o);
                        key = t < key ? t : key;
                    }
                    if (NWARPS > 1) {
                        __syncthreads();
                        if (lane == 0) s.pk[warp] = key;
                        __syncthreads();
                        key = s.pk[0];
#pragma unroll
                        for (int w = 1; w < NWARPS; ++w) key = s.pk[w] < key ? s.pk[w] : key;
                    }
                    jstar = static_cast<int>(key & 0xffffffffu);
                }
                const int pos = replay_pos(jstar, R, P, s.rm_idx);
                const int r4c = s.row4col[jstar];
                if (tid == 0) {
                    s.rm_col[R] = jstar;
                    s.rm_idx[R] = pos;
                }
                if ((jstar % T) == tid) s.sc[jstar] = 1;
                ++R;
                if (r4c < 0) break;   // jstar is the sink
                i = r4c;
                if (NWARPS == 1) __syncwarp();
            }
            __syncthreads();
            if (failed) break;
            // ---- dual update (u, v) over the scanned rows / columns
            for (int k = tid; k < R; k += T) {
                const int j = s.rm_col[k];
                const double delta = __dsub_rn(min_val, s.spc[j]);
                s.v[j] = __dsub_rn(s.v[j], delta);
                if (k < R - 1) {
                    const int row = s.row4col[j];
                    s.u[row] = __dadd_rn(s.u[row], delta);
                }
                s.sc[j] = 0;
            }
            if (tid == 0) s.u[cur] = __dadd_rn(s.u[cur], min_val);
            __syncthreads();
            // ---- augment along the path
            if (tid == 0) {
                int j = s.rm_col[R - 1];
                for (;;) {
                    const int r = s.path[j];
                    s.row4col[j] = static_cast<short>(r);
                    const int t = s.col4row[r];
                    s.col4row[r] = j;
                    j = t;
                    if (r == cur) break;
                }
            }
            __syncthreads();
        }

        // ---- epilogue: mask, matched GT index, loss terms, gradients
        double acc_sq = 0.0, acc_conf = 0.0;
        int n_match = 0;
        for (int j = tid; j < P; j += T) {
            const int r = s.row4col[j];
            if (p.mask) p.mask[row0 + j] = r >= 0 ? 1 : 0;
            if (p.gt_idx) p.gt_idx[row0 + j] = r;
            n_match += r >= 0;
            {
                const float c = boundary ? s.conf[j] : __fadd_rn(s.conf[j], kEps32);
                float4 dl = make_float4(0.f, 0.f, 0.f, 0.f);
                float dc;
                if (r >= 0) {
                    const float4 l = s.loc[j], g = s.gt[r];
                    const float d0 = __fsub_rn(l.x, g.x), d1 = __fsub_rn(l.y, g.y), d2 = __fsub_rn(l.z, g.z),
                                d3 = __fsub_rn(l.w, g.w);
                    acc_sq += static_cast<double>(__fmul_rn(d0, d0));
                    acc_sq += static_cast<double>(__fmul_rn(d1, d1));
                    acc_sq += static_cast<double>(__fmul_rn(d2, d2));
                    acc_sq += static_cast<double>(__fmul_rn(d3, d3));
                    dl = make_float4(__fmul_rn(p.alpha, d0), __fmul_rn(p.alpha, d1), __fmul_rn(p.alpha, d2),
                                     __fmul_rn(p.alpha, d3));
                    acc_conf -= static_cast<double>(s.lc[j]);
                    dc = __fdiv_rn(-1.0f, c);
                } else {
                    const float one_m = __fsub_rn(1.0f, c);
                    const float arg = __fadd_rn(one_m, kEps32);
                    float vcl = one_m;
                    if (vcl > 1.0f) vcl = 1.0f;
                    if (vcl <= 0.0f) vcl = kEps32;
                    const float la = (arg == vcl) ? s.l1[j] : nplogf(arg);
                    acc_conf -= static_cast<double>(la);
                    dc = __fdiv_rn(1.0f, arg);
                }
                if (logits) {
                    const float s0 = s.conf[j];   // d sigmoid / d logit = s (1 - s)
                    dc = __fmul_rn(dc, __fmul_rn(s0, __fsub_rn(1.0f, s0)));
                }
                if (p.d_loc) st_stream_f4(reinterpret_cast<float4 *>(p.d_loc) + row0 + j, dl);
                if (p.d_conf) p.d_conf[row0 + j] = dc;
            }
        }
        // stacked GT rows in ascending prior order: rank by counting (n <= M is small)
        if (p.stacked && !failed) {
            const int off = p.stk_offsets[b];
            for (int i = tid; i < n; i += T) {
                const int pi = s.col4row[i];
                int rank = 0;
                for (int q = 0; q < n; ++q) rank += s.col4row[q] < pi;
                reinterpret_cast<float4 *>(p.stacked)[off + rank] = s.gt[i];
            }
        }
        {
            acc_sq = warp_sum(acc_sq);
            acc_conf = warp_sum(acc_conf);
            n_match = __reduce_add_sync(0xffffffffu, n_match);
            __syncthreads();   // s.red reuse across images
            if (lane == 0) {
                s.red[warp] = acc_sq;
                s.red[NWARPS + warp] = acc_conf;
                s.ri[warp] = n_match;
            }
            __syncthreads();
            if (tid == 0) {
                double a = 0.0, c = 0.0;
                int m = 0;
                for (int w = 0; w < NWARPS; ++w) {
                    a += s.red[w];
                    c += s.red[NWARPS + w];
                    m += s.ri[w];
                }
                p.partials[2 * b] = a;
                p.partials[2 * b + 1] = c;
                p.img_matched[b] = m;
            }
        }
        __syncthreads();   // shared state is reused by the next image
    }

    if (status) atomicOr(p.status, status);

    // ---- last CTA to finish reduces the per-image partials in a fixed order
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned t = atomicAdd(p.ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const TailPrefetch pre = tail_prefetch(p);
    double a = 0.0, c = 0.0;
    long long m = 0;
    for (int b = tid; b < p.B; b += T) {
        a += __ldcg(p.partials + 2 * b);
        c += __ldcg(p.partials + 2 * b + 1);
        m += __ldcg(p.img_matched + b);
    }
    a = warp_sum(a);
    c = warp_sum(c);
    double md = warp_sum(static_cast<double>(m));
    if (lane == 0) {
        s.red[warp] = a;
        s.red[NWARPS + warp] = c;
        s.pv[warp] = md;
    }
    __syncthreads();
    if (warp == 0) {
        double A = 0.0, C = 0.0, Mt = 0.0;
        for (int w = 0; w < NWARPS; ++w) {
            A += s.red[w];
            C += s.red[NWARPS + w];
            Mt += s.pv[w];
        }
        finalize_losses(p, A, C, Mt, pre);
    }
}

struct WsLayout {
    size_t partials, img_matched, offsets, order, ticket, status, ar_seq, lseq, sched, total;
};
static WsLayout ws_layout(int B) {
    WsLayout w;
    size_t o = 0;
    w.ticket = o;
    o += 8;
    w.status = o;
    o += 8;
    w.ar_seq = o;
    o += 8;
    w.lseq = o;      // launch sequence number published with the results
    o += 8;
    w.sched = o;     // dynamic image scheduler: two slots {queue counter, order-ready flag}
    o += 16;
    w.partials = o;
    o += sizeof(double) * 2 * static_cast<size_t>(B);
    w.img_matched = o;
    o += sizeof(int32_t) * static_cast<size_t>(B);
    o = align_up(o, 16);
    w.offsets = o;
    o += sizeof(int32_t) * (static_cast<size_t>(B) + 1);
    o = align_up(o, 16);
    w.order = o;     // two heavy-first orders (alternating launches)
    o += sizeof(int32_t) * static_cast<size_t>(B) * 2;
    w.total = align_up(o, 256);
    return w;
}

template <int NWARPS>
static int launch_match(const MatchParams &p, size_t smem, int grid, cudaStream_t st) {
    auto kern = mbx_match_loss_kernel<NWARPS>;
    static thread_local size_t configured = 0;
    if (smem > configured) {
        if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    static_cast<int>(smem)),
                               "cudaFuncSetAttribute(match)"))
            return e;
        configured = smem;
    }
    kern<<<grid, NWARPS * 32, smem, st>>>(p);
    return check_cuda(cudaGetLastError(), "launch mbx_match_loss_kernel");
}

template <int NWARPS>
static int occupancy(size_t smem) {
    int nb = 0;
    cudaFuncSetAttribute(mbx_match_loss_kernel<NWARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(smem));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, mbx_match_loss_kernel<NWARPS>, NWARPS * 32, smem);
    return nb;
}


// ---- diagnostics -----------------------------------------------------------
__global__ void mbx_debug_nplog_kernel(const float *in, float *out, long long n) {
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += 256ll * gridDim.x) out[i] = nplogf(in[i]);
}

__global__ void mbx_debug_sqrt_kernel(unsigned first, unsigned count, unsigned long long *mismatches) {
    unsigned long long bad = 0;
    for (unsigned long long k = blockIdx.x * 256ull + threadIdx.x; k < count; k += 256ull * gridDim.x) {
        const float x = __uint_as_float(first + static_cast<unsigned>(k));
        const float a = sqrt_rn_branchfree(x), b = __fsqrt_rn(x);
        const bool same = (__float_as_uint(a) == __float_as_uint(b)) || (a != a && b != b);
        bad += same ? 0 : 1;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// |__logf(x) - numpy log(x)| against the bound mbx_bound.h assumes for the cheap cost form
__global__ void mbx_debug_fastlog_kernel(unsigned first, unsigned count, unsigned long long *violations) {
    unsigned long long bad = 0;
    for (unsigned long long k = blockIdx.x * 256ull + threadIdx.x; k < count; k += 256ull * gridDim.x) {
        const float x = __uint_as_float(first + static_cast<unsigned>(k));
        const float a = __logf(x), b = nplogf(x);
        bool ok;
        if (b != b)
            ok = a != a;
        else if (b == CUDART_INF_F || b == -CUDART_INF_F)
            ok = a == b;
        else
            ok = fabs(static_cast<double>(a) - static_cast<double>(b)) <=
                 4.76837158203125e-07 + 1.9073486328125e-06 * fabs(static_cast<double>(a));   // 2^-21 + 2^-19 |a|
        bad += ok ? 0 : 1;
    }
    if (bad) atomicAdd(violations, bad);
}

__global__ void mbx_debug_cost_kernel(const float *loc, const float *conf, const float *gt, int P, int n,
                                      float alpha, double *C) {
    const float half_alpha = __fdiv_rn(alpha, 2.0f);
    for (int p = blockIdx.x * 256 + threadIdx.x; p < P; p += 256 * gridDim.x) {
        const float c = conf[p];
        const float lc = nplogf(c);
        float v = __fsub_rn(1.0f, c);
        if (v > 1.0f) v = 1.0f;
        if (v <= 0.0f) v = kEps32;
        const float l1 = nplogf(v);
        const float4 l = reinterpret_cast<const float4 *>(loc)[p];
        for (int j = 0; j < n; ++j)
            C[static_cast<size_t>(p) * n + j] =
                static_cast<double>(cost32(l, reinterpret_cast<const float4 *>(gt)[j], half_alpha, lc, l1));
    }
}

}  // namespace mbx

using namespace mbx;

extern "C" int mbx_debug_nplog(const float *in, float *out, long long n, void *stream) {
    if (n <= 0) return 0;
    long long blocks = (n + 255) / 256;
    int grid = static_cast<int>(blocks < 148 * 16 ? blocks : 148 * 16);
    mbx_debug_nplog_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, n);
    return check_cuda(cudaGetLastError(), "launch mbx_debug_nplog_kernel");
}

extern "C" int mbx_debug_sqrt_mismatches(unsigned first_bits, unsigned count, unsigned long long *mismatches,
                                         void *stream) {
    if (!mismatches) return MBX_E_ARG;
    mbx_debug_sqrt_kernel<<<148 * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(first_bits, count, mismatches);
    return check_cuda(cudaGetLastError(), "launch mbx_debug_sqrt_kernel");
}

extern "C" int mbx_debug_fastlog_violations(unsigned first_bits, unsigned count, unsigned long long *violations,
                                            void *stream) {
    if (!violations) return MBX_E_ARG;
    mbx_debug_fastlog_kernel<<<148 * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(first_bits, count, violations);
    return check_cuda(cudaGetLastError(), "launch mbx_debug_fastlog_kernel");
}

extern "C" int mbx_debug_cost_matrix(const float *loc, const float *conf, const float *gt, int P, int n,
                                     float alpha, double *C, void *stream) {
    if (P <= 0 || n <= 0) return 0;
    mbx_debug_cost_kernel<<<(P + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(loc, conf, gt, P, n, alpha, C);
    return check_cuda(cudaGetLastError(), "launch mbx_debug_cost_kernel");
}

extern "C" size_t mbx_match_workspace_bytes(int B, int P, int M) {
    (void)P;
    (void)M;
    return ws_layout(B < 1 ? 1 : B).total;
}

namespace mbx {
int match_loss_impl(const mbx_heads *heads, const float *locations, const float *confidences,
                           const float *gt_bboxes, const int32_t *num_gt, const int32_t *gt_row,
                           const float *priors, int B, int P, int M, float alpha, unsigned flags, int32_t *mask,
                           int32_t *matched_gt_idx, float *stacked_gt, int32_t *n_stacked, float *d_locations,
                           float *d_confidences, float *confidences_out, float *results, void *workspace,
                           size_t workspace_bytes, const unsigned long long *peer_buffers, int world, int rank,
                           void *stream);
}

extern "C" int mbx_match_loss(const float *locations, const float *confidences, const float *gt_bboxes,
                              const int32_t *num_gt, const float *priors, int B, int P, int M, float alpha,
                              unsigned flags, int32_t *mask, int32_t *matched_gt_idx, float *stacked_gt,
                              int32_t *n_stacked, float *d_locations, float *d_confidences,
                              float *confidences_out, float *results, void *workspace, size_t workspace_bytes,
                              void *stream) {
    return match_loss_impl(nullptr, locations, confidences, gt_bboxes, num_gt, nullptr, priors, B, P, M, alpha, flags,
                           mask, matched_gt_idx, stacked_gt, n_stacked, d_locations, d_confidences, confidences_out,
                           results, workspace, workspace_bytes, nullptr, 1, 0, stream);
}

extern "C" int mbx_match_loss_ragged(const float *locations, const float *confidences, const float *gt_flat,
                                     const int32_t *gt_row_offsets, const float *priors, int B, int P, int M,
                                     float alpha, unsigned flags, int32_t *mask, int32_t *matched_gt_idx,
                                     float *stacked_gt, int32_t *n_stacked, float *d_locations,
                                     float *d_confidences, float *confidences_out, float *results,
                                     void *workspace, size_t workspace_bytes, void *stream) {
    if (!gt_row_offsets) {
        set_error("mbx_match_loss_ragged: null gt_row_offsets");
        return MBX_E_ARG;
    }
    return match_loss_impl(nullptr, locations, confidences, gt_flat, nullptr, gt_row_offsets, priors, B, P, M, alpha,
                           flags, mask, matched_gt_idx, stacked_gt, n_stacked, d_locations, d_confidences,
                           confidences_out, results, workspace, workspace_bytes, nullptr, 1, 0, stream);
}

extern "C" int mbx_match_loss_heads(const mbx_heads *heads, const float *gt_bboxes, const int32_t *num_gt,
                                    const int32_t *gt_row_offsets, const float *priors, int B, int P, int M,
                                    float alpha, unsigned flags, int32_t *mask, int32_t *matched_gt_idx,
                                    float *stacked_gt, int32_t *n_stacked, float *confidences_out, float *results,
                                    void *workspace, size_t workspace_bytes, void *stream) {
    if (!heads) {
        set_error("mbx_match_loss_heads: null heads");
        return MBX_E_ARG;
    }
    return match_loss_impl(heads, nullptr, nullptr, gt_bboxes, num_gt, gt_row_offsets, priors, B, P, M, alpha, flags,
                           mask, matched_gt_idx, stacked_gt, n_stacked, nullptr, nullptr, confidences_out, results,
                           workspace, workspace_bytes, nullptr, 1, 0, stream);
}

extern "C" size_t mbx_allreduce_buffer_bytes(void) { return align_up(kArBytes, 256); }

extern "C" int mbx_allreduce_config(int pdl_lag, int relay_batch);

namespace mbx {
__global__ void mbx_allreduce_flush_kernel(MatchParams p) {   // one warp
    const unsigned seq = *p.ar_seq;
    if (seq == 0u) return;
    double g_loc = 0.0, g_conf = 0.0;
    const bool ok = ar_pull_warp(p, seq - 1u, g_loc, g_conf);   // every rank's newest step, from its outbox
    if (threadIdx.x != 0) return;
    if (ok) {
        double *r64 = reinterpret_cast<double *>(p.results);
        r64[4] = g_loc;
        r64[5] = g_conf;
        p.results[12] = static_cast<float>(g_loc);
        p.results[13] = static_cast<float>(g_conf);
        p.results[14] = static_cast<float>(seq - 1u);
    } else {
        p.results[2] = static_cast<float>(static_cast<unsigned>(p.results[2]) | MBX_STATUS_AR_TIMEOUT);
    }
}

// The RELAY of the deferred fused all-reduce (see mbx_match.cuh): one warp on a side stream.  Polls this rank's
// outbox (local memory) for the next step's words and forwards them into every rank's table (lane r: four
// 8-byte NVLink stores into rank r's buffer; lane == rank: the own table).  Forwards at most `max_steps` steps,
// exits when no new step appears for `idle_cycles` (the host side launches the next relay kRelaySteps launches
// later) or the sticky timeout flag is up.  Steps at least kArRing / 2 behind this rank's step counter are
// skipped: every rank has consumed them (ranks are never more than ar_lag steps apart).
__global__ void mbx_allreduce_relay_kernel(MatchParams p, int max_steps, long long idle_cycles, int batch) {
    const int lane = threadIdx.x & 31;
    unsigned char *mine = reinterpret_cast<unsigned char *>(p.ar_peer[p.ar_rank]);
    volatile unsigned *relayed = reinterpret_cast<volatile unsigned *>(mine + kArRelayedOffset);
    volatile unsigned *seqp = reinterpret_cast<volatile unsigned *>(p.ar_seq);
    volatile unsigned *dead = ar_dead_flag(p);
    unsigned s = *relayed;
    long long t0 = clock64();
    for (int n = 0; n < max_steps;) {
        const unsigned seq = *seqp;
        if (static_cast<int>(seq - s) >= kArRing / 2) s = seq - kArRing / 2 + 1u;   // (stale: consumed everywhere)
        // the NEWEST step of the next batch: when its words are there, the older ones of the batch are, too
        // (a rank completes its steps in order)
        const unsigned hi = s + static_cast<unsigned>(batch) - 1u;
        const unsigned long long *wh = ar_outbox(p.ar_peer[p.ar_rank], hi);
        const unsigned long long h0 = ld_relaxed_sys_u64(wh), h3 = ld_relaxed_sys_u64(wh + 3);
        const bool idle = clock64() - t0 > idle_cycles;
        if ((static_cast<unsigned>(h0 >> 32) == hi + 1u && static_cast<unsigned>(h3 >> 32) == hi + 1u) || idle) {
            // (idle: forward whatever part of the batch exists, then leave)
            int sent = 0;
            for (unsigned q = s; q <= hi; ++q) {
                const unsigned long long *w = ar_outbox(p.ar_peer[p.ar_rank], q);
                const unsigned long long w0 = ld_relaxed_sys_u64(w), w1 = ld_relaxed_sys_u64(w + 1),
                                         w2 = ld_relaxed_sys_u64(w + 2), w3 = ld_relaxed_sys_u64(w + 3);
                const unsigned tag = q + 1u;
                if (!(static_cast<unsigned>(w0 >> 32) == tag && static_cast<unsigned>(w1 >> 32) == tag &&
                      static_cast<unsigned>(w2 >> 32) == tag && static_cast<unsigned>(w3 >> 32) == tag))
                    break;
                if (lane < p.ar_world) {
                    unsigned long long *d = ar_words(p.ar_peer[lane], q, p.ar_rank);
                    st_relaxed_sys_u64(d, w0);
                    st_relaxed_sys_u64(d + 1, w1);
                    st_relaxed_sys_u64(d + 2, w2);
                    st_relaxed_sys_u64(d + 3, w3);
                }
                ++sent;
            }
            s += sent;
            n += sent;
            if (lane == 0 && sent) *relayed = s;
            if (idle) break;
            t0 = clock64();
            continue;
        }
        if (*dead != 0u) break;
    }
}
}  // namespace mbx

namespace {
int g_pdl_lag = 12, g_relay_batch = kRelaySteps;
// Host side of the relay: one side stream per device and a launch counter per symmetric buffer (thread-local,
// like the scheduler's launch ids).  Every kRelaySteps-th deferred launch on a buffer enqueues a relay for the
// next kRelaySteps steps BEFORE the step itself; relays of one device run one after the other on the side
// stream, so a relay that finds its steps already forwarded moves on at once.  Not during stream capture
// (a captured step has no host side when it is replayed: its reductions take the pull route).
struct RelayHost {
    int dev = -1;
    cudaStream_t side = nullptr;
    std::unordered_map<unsigned long long, unsigned> launches;
};
void maybe_launch_relay(const MatchParams &p, cudaStream_t st) {
    static thread_local RelayHost host;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != host.dev) {   // (one side stream per device; a process of this framework drives one GPU)
        host.side = nullptr;
        host.launches.clear();
        host.dev = dev;
    }
    unsigned &n = host.launches[p.ar_peer[p.ar_rank]];
    if (n++ % static_cast<unsigned>(kRelaySteps) != 0u) return;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
        cudaGetLastError();
        return;
    }
    if (!host.side && cudaStreamCreateWithFlags(&host.side, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        host.side = nullptr;
        return;   // (no relay: the pull route still completes every reduction)
    }
    // idle limit: ~30 us at 2 GHz -- a relay outlives the gaps between back-to-back steps of the latency-bound
    // shapes it exists for, and holds a trailing synchronisation back by no more than that
    // (without PDL a step completes the PREVIOUS step's reduction: its sums must be forwarded one by one)
    mbx_allreduce_relay_kernel<<<1, 32, 0, host.side>>>(p, kRelaySteps, 60000ll,
                                                       (p.flags & MBX_FLAG_PDL) ? g_relay_batch : 1);
    cudaGetLastError();
}
}  // namespace

extern "C" int mbx_allreduce_config(int pdl_lag, int relay_batch) {
    if (pdl_lag < 1 || relay_batch < 1 || relay_batch > kRelaySteps || 2 * pdl_lag + relay_batch >= kArRing) {
        set_error("mbx_allreduce_config: need 1 <= relay_batch <= %d and 2 * pdl_lag + relay_batch < %d", kRelaySteps,
                  kArRing);
        return MBX_E_ARG;
    }
    g_pdl_lag = pdl_lag;
    g_relay_batch = relay_batch;
    return 0;
}

extern "C" int mbx_allreduce_flush(float *results, void *workspace, size_t workspace_bytes,
                                   const unsigned long long *peer_buffers, int world, int rank, void *stream) {
    if (world < 2 || world > MBX_MAX_PEERS || rank < 0 || rank >= world || !peer_buffers || !results || !workspace) {
        set_error("mbx_allreduce_flush: bad arguments");
        return MBX_E_ARG;
    }
    (void)workspace_bytes;
    MatchParams p{};
    p.results = results;
    p.ar_seq = reinterpret_cast<unsigned *>(peer_buffers[rank] + kArSeqOffset);
    p.ar_world = world;
    p.ar_rank = rank;
    for (int r = 0; r < MBX_MAX_PEERS; ++r) p.ar_peer[r] = r < world ? peer_buffers[r] : 0ull;
    mbx_allreduce_flush_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
    return check_cuda(cudaGetLastError(), "launch mbx_allreduce_flush_kernel");
}

extern "C" int mbx_match_loss_allreduce(const float *locations, const float *confidences, const float *gt_bboxes,
                                        const int32_t *num_gt, const float *priors, int B, int P, int M,
                                        float alpha, unsigned flags, int32_t *mask, int32_t *matched_gt_idx,
                                        float *stacked_gt, int32_t *n_stacked, float *d_locations,
                                        float *d_confidences, float *confidences_out, float *results,
                                        void *workspace, size_t workspace_bytes,
                                        const unsigned long long *peer_buffers, int world, int rank,
                                        void *stream) {
    return match_loss_impl(nullptr, locations, confidences, gt_bboxes, num_gt, nullptr, priors, B, P, M, alpha, flags,
                           mask, matched_gt_idx, stacked_gt, n_stacked, d_locations, d_confidences, confidences_out,
                           results, workspace, workspace_bytes, peer_buffers, world, rank, stream);
}

// ---- prepared launches ------------------------------------------------------
// A plan is the argument list of mbx_match_loss_allreduce, validated once and kept in a caller-owned
// object: launching it costs one foreign call with two arguments (for latency-critical training loops
// whose steps are shorter than the host's argument marshalling).
struct mbx_match_plan {
    const float *locations, *confidences, *gt_bboxes, *priors;
    const int32_t *num_gt;
    int B, P, M;
    float alpha;
    unsigned flags;
    int32_t *mask, *matched_gt_idx, *n_stacked;
    float *stacked_gt, *d_locations, *d_confidences, *confidences_out, *results;
    void *workspace;
    size_t workspace_bytes;
    unsigned long long peers[MBX_MAX_PEERS];
    int world, rank;
};

extern "C" int mbx_match_plan_create(mbx_match_plan **plan, const float *locations, const float *confidences,
                                     const float *gt_bboxes, const int32_t *num_gt, const float *priors, int B, int P,
                                     int M, float alpha, unsigned flags, int32_t *mask, int32_t *matched_gt_idx,
                                     float *stacked_gt, int32_t *n_stacked, float *d_locations, float *d_confidences,
                                     float *confidences_out, float *results, void *workspace, size_t workspace_bytes,
                                     const unsigned long long *peer_buffers, int world, int rank) {
    if (!plan) {
        set_error("mbx_match_plan_create: null plan pointer");
        return MBX_E_ARG;
    }
    *plan = nullptr;
    if (world < 1 || world > MBX_MAX_PEERS || rank < 0 || rank >= world || (world > 1 && !peer_buffers)) {
        set_error("mbx_match_plan_create: bad world=%d rank=%d (max %d ranks)", world, rank, MBX_MAX_PEERS);
        return MBX_E_ARG;
    }
    mbx_match_plan *pl = new (std::nothrow) mbx_match_plan{locations, confidences, gt_bboxes, priors, num_gt, B, P, M,
                                                          alpha, flags, mask, matched_gt_idx, n_stacked, stacked_gt,
                                                          d_locations, d_confidences, confidences_out, results,
                                                          workspace, workspace_bytes, {0}, world, rank};
    if (!pl) {
        set_error("mbx_match_plan_create: out of host memory");
        return MBX_E_ARG;
    }
    for (int r = 0; r < MBX_MAX_PEERS; ++r) pl->peers[r] = (world > 1 && r < world) ? peer_buffers[r] : 0ull;
    *plan = pl;
    return 0;
}

extern "C" int mbx_match_plan_launch(const mbx_match_plan *pl, void *stream) {
    if (!pl) {
        set_error("mbx_match_plan_launch: null plan");
        return MBX_E_ARG;
    }
    return match_loss_impl(nullptr, pl->locations, pl->confidences, pl->gt_bboxes, pl->num_gt, nullptr, pl->priors, pl->B,
                           pl->P, pl->M, pl->alpha, pl->flags, pl->mask, pl->matched_gt_idx, pl->stacked_gt, pl->n_stacked,
                           pl->d_locations, pl->d_confidences, pl->confidences_out, pl->results, pl->workspace,
                           pl->workspace_bytes, pl->world > 1 ? pl->peers : nullptr, pl->world, pl->rank, stream);
}

extern "C" int mbx_match_plan_launch_staged(const mbx_match_plan *pl, const void *host_src, void *dev_dst,
                                            size_t nbytes, void *stream) {
    if (!pl || (nbytes > 0 && (!host_src || !dev_dst))) {
        set_error("mbx_match_plan_launch_staged: null plan / buffer");
        return MBX_E_ARG;
    }
    if (nbytes > 0) {
        if (int e = check_cuda(cudaMemcpyAsync(dev_dst, host_src, nbytes, cudaMemcpyHostToDevice,
                                               static_cast<cudaStream_t>(stream)),
                               "cudaMemcpyAsync(staged inputs)"))
            return e;
    }
    return mbx_match_plan_launch(pl, stream);
}

extern "C" void mbx_match_plan_destroy(mbx_match_plan *pl) { delete pl; }

int mbx::match_loss_impl(const mbx_heads *heads, const float *locations, const float *confidences,
                                const float *gt_bboxes, const int32_t *num_gt, const int32_t *gt_row,
                                const float *priors, int B, int P, int M, float alpha, unsigned flags,
                                int32_t *mask, int32_t *matched_gt_idx, float *stacked_gt, int32_t *n_stacked,
                                float *d_locations, float *d_confidences, float *confidences_out, float *results,
                                void *workspace, size_t workspace_bytes, const unsigned long long *peer_buffers,
                                int world, int rank, void *stream) {
    if (world < 1 || world > MBX_MAX_PEERS || rank < 0 || rank >= world || (world > 1 && !peer_buffers)) {
        set_error("mbx_match_loss_allreduce: bad world=%d rank=%d (max %d ranks)", world, rank, MBX_MAX_PEERS);
        return MBX_E_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (B < 0 || P <= 0 || M < 0) {
        set_error("mbx_match_loss: bad sizes B=%d P=%d M=%d", B, P, M);
        return MBX_E_ARG;
    }
    if (B == 0) return 0;
    const bool boundary = (flags & MBX_FLAG_BOUNDARY) != 0;
    if (boundary && (flags & MBX_FLAG_LOGITS)) {
        set_error("mbx_match_loss: MBX_FLAG_BOUNDARY and MBX_FLAG_LOGITS are exclusive");
        return MBX_E_ARG;
    }
    if ((!heads && (!locations || !confidences)) || (!num_gt && !gt_row) || (M > 0 && !gt_bboxes && !gt_row) ||
        (!boundary && !priors) || !workspace || !results) {
        set_error("mbx_match_loss: null input / results / workspace pointer");
        return MBX_E_ARG;
    }
    if (heads) {
        int tot = 0;
        bool bad = heads->num_heads < 1 || heads->num_heads > MBX_MAX_HEADS || boundary;
        for (int k = 0; !bad && k < heads->num_heads; ++k) {
            bad = heads->head_priors[k] <= 0 || !heads->locations[k] || !heads->confidences[k] ||
                  (reinterpret_cast<uintptr_t>(heads->locations[k]) & 15u) ||
                  (reinterpret_cast<uintptr_t>(heads->d_locations[k]) & 15u) ||
                  ((heads->d_locations[k] != nullptr) != (heads->d_locations[0] != nullptr)) ||
                  ((heads->d_confidences[k] != nullptr) != (heads->d_confidences[0] != nullptr));
            tot += heads->head_priors[k];
        }
        if (bad || tot != P) {
            set_error("mbx_match_loss_heads: bad head table (num_heads=%d, sum of head_priors=%d, P=%d; no "
                      "MBX_FLAG_BOUNDARY; pointers 16-byte aligned; gradients for all heads or none)",
                      heads->num_heads, tot, P);
            return MBX_E_ARG;
        }
        if (flags & MBX_FLAG_GENERIC) {
            set_error("mbx_match_loss_heads: the per-head layout needs the register-resident kernel family");
            return MBX_E_ARG;
        }
    }
    auto mis16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15u) != 0; };
    if (mis16(locations) || mis16(gt_bboxes) || mis16(priors) || mis16(stacked_gt) || mis16(d_locations) ||
        mis16(results) || mis16(workspace)) {
        set_error("mbx_match_loss: pointers must be 16-byte aligned");
        return MBX_E_ARG;
    }
    if (M > P || M > 32767 || P >= kNoCol) {
        set_error("mbx_match_loss: need M <= P and M <= 32767 (got P=%d M=%d)", P, M);
        return MBX_E_TOO_LARGE;
    }
    const WsLayout wl = ws_layout(B);
    if (workspace_bytes < wl.total) {
        set_error("mbx_match_loss: workspace %zu < %zu bytes", workspace_bytes, wl.total);
        return MBX_E_WORKSPACE;
    }
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    MatchParams p;
    p.locations = locations;
    p.confidences = confidences;
    p.gt = gt_bboxes;
    p.priors = priors;
    p.num_gt = num_gt;
    p.gt_row = gt_row;
    p.nheads = 1;
    for (int k = 0; k < MBX_MAX_HEADS; ++k) p.heads[k] = HeadTab{nullptr, nullptr, nullptr, nullptr, 0, 0};
    if (heads) {
        int off = 0;
        p.nheads = heads->num_heads;
        for (int k = 0; k < heads->num_heads; ++k) {
            p.heads[k] = HeadTab{heads->locations[k], heads->confidences[k], heads->d_locations[k],
                                 heads->d_confidences[k], heads->head_priors[k], off};
            off += heads->head_priors[k];
        }
        if (p.nheads == 1) {   // a single head is the dense layout
            p.locations = heads->locations[0];
            p.confidences = heads->confidences[0];
            d_locations = heads->d_locations[0];
            d_confidences = heads->d_confidences[0];
        }
    }
    p.B = B;
    p.P = P;
    p.M = M;
    p.alpha = alpha;
    p.flags = flags;
    p.mask = mask;
    p.gt_idx = matched_gt_idx;
    p.stacked = stacked_gt;
    p.n_stacked = n_stacked;
    p.d_loc = d_locations;
    p.d_conf = d_confidences;
    p.conf_out = confidences_out;
    p.results = results;
    p.partials = reinterpret_cast<double *>(ws + wl.partials);
    p.img_matched = reinterpret_cast<int32_t *>(ws + wl.img_matched);
    p.stk_offsets = reinterpret_cast<int32_t *>(ws + wl.offsets);
    p.ticket = reinterpret_cast<unsigned *>(ws + wl.ticket);
    p.order_base = reinterpret_cast<int32_t *>(ws + wl.order);
    p.sched_base = reinterpret_cast<unsigned *>(ws + wl.sched);
    p.order = p.order_base;
    p.queue = p.sched_base;
    p.oready = p.sched_base + 2;
    p.launch_id = 1u;
    p.order_first = 0;
    p.lseq = reinterpret_cast<unsigned *>(ws + wl.lseq);
    p.dynamic = 0;
    p.status = reinterpret_cast<unsigned *>(ws + wl.status);
    // the step counter lives in this rank's own symmetric buffer, next to the arrival counters it
    // must stay consistent with (a re-allocated workspace must not reset it)
    p.ar_seq = world > 1 ? reinterpret_cast<unsigned *>(peer_buffers[rank] + kArSeqOffset)
                         : reinterpret_cast<unsigned *>(ws + wl.ar_seq);
    p.ar_world = world;
    p.ar_rank = rank;
    p.ar_lag_pdl = static_cast<unsigned>(g_pdl_lag);
    for (int r = 0; r < MBX_MAX_PEERS; ++r) p.ar_peer[r] = (world > 1 && r < world) ? peer_buffers[r] : 0ull;

    int nwarps = static_cast<int>((flags >> MBX_FLAG_WARPS_SHIFT) & 0xffu);
    const int ncols = static_cast<int>((flags >> MBX_FLAG_COLS_SHIFT) & 0xffu);
    if (stacked_gt || n_stacked) {
        p.flags &= ~MBX_FLAG_PDL;   // the scan kernel right before the matching kernel produces one of its inputs
        mbx_scan_num_gt_kernel<<<1, 1024, 0, st>>>(num_gt, gt_row, B, M, p.stk_offsets, n_stacked);
        if (int e = check_cuda(cudaGetLastError(), "launch mbx_scan_num_gt_kernel")) return e;
    }
    if (flags & MBX_FLAG_GENERIC) p.flags &= ~MBX_FLAG_PDL;   // (the generic kernel has no early-start path)
    if (world > 1 && (flags & MBX_FLAG_AR_DEFERRED)) maybe_launch_relay(p, st);
    if (!(flags & MBX_FLAG_GENERIC)) {
        // register-resident family first; it declines shapes it has no instantiation for
        const int rc = launch_match_reg(p, nwarps, ncols, st);
        if (rc != MBX_E_TOO_LARGE) return rc;
        if (p.nheads > 1) {
            set_error("mbx_match_loss_heads: P=%d is beyond the register-resident kernel family", P);
            return rc;
        }
        if (nwarps > 8) nwarps = 8;
    }
    if (nwarps == 0) nwarps = P <= 256 ? 2 : (P <= 1024 ? 4 : 8);
    if (nwarps != 1 && nwarps != 2 && nwarps != 4 && nwarps != 8) {
        set_error("mbx_match_loss: forced warps must be 1, 2, 4 or 8");
        return MBX_E_ARG;
    }
    const size_t smem = carve(nullptr, nullptr, P, M, nwarps, !boundary);
    if (smem > static_cast<size_t>(max_smem_optin())) {
        set_error("mbx_match_loss: P=%d M=%d needs %zu bytes of shared memory per CTA (max %d)", P, M, smem,
                  max_smem_optin());
        return MBX_E_TOO_LARGE;
    }
    int occ = 1;
    switch (nwarps) {
        case 1: occ = occupancy<1>(smem); break;
        case 2: occ = occupancy<2>(smem); break;
        case 4: occ = occupancy<4>(smem); break;
        default: occ = occupancy<8>(smem); break;
    }
    if (occ < 1) occ = 1;
    int grid = sm_count() * occ;
    if (grid > B) grid = B;
    int rc;
    switch (nwarps) {
        case 1: rc = launch_match<1>(p, smem, grid, st); break;
        case 2: rc = launch_match<2>(p, smem, grid, st); break;
        case 4: rc = launch_match<4>(p, smem, grid, st); break;
        default: rc = launch_match<8>(p, smem, grid, st); break;
    }
    return rc;
}
