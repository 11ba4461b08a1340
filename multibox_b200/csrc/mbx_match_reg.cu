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
#include "mbx_match.cuh"

namespace mbx {

namespace {

constexpr unsigned kPayNone = 0xffffffffu;
constexpr unsigned kPayTie = 1u << 15;

struct RSmem {
    float4 *priors, *gt;
    double *u, *red;
    int4 *part;                 // [2][NWARPS] {key_hi, key_lo, payload, -}
    unsigned long long *pk;     // [NWARPS]
    int *col4row, *rm_col, *rm_idx, *rm_pm, *visit, *ri;
    short *row4col;
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
    size_t o_r4c = take(sizeof(short) * P, 2);
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

template <int NWARPS, int C>
__global__ void __launch_bounds__(NWARPS * 32) mbx_match_loss_reg_kernel(const MatchParams p) {
    constexpr int T = NWARPS * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RSmem s;
    const bool boundary = (p.flags & MBX_FLAG_BOUNDARY) != 0;
    const bool logits = (p.flags & MBX_FLAG_LOGITS) != 0;
    const bool has_priors = !boundary;
    rcarve(&s, smem_raw, p.P, p.M, NWARPS, has_priors);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = p.P, M = p.M;
    const float half_alpha = __fdiv_rn(p.alpha, 2.0f);   // (alpha / 2.) in fp32, loss.py:35
    const double INF = CUDART_INF;
    unsigned status = 0;

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
    }
    bool priors_ready = !has_priors;
    int pbuf = 0;

    unsigned invalid_mask = 0;   // columns of this thread beyond P
#pragma unroll
    for (int c = 0; c < C; ++c)
        if (tid + c * T >= P) invalid_mask |= 1u << c;

    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        int n = p.num_gt[b];
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
        double v[C], spc[C];
        int r4c[C], pm[C];
        const float4 *gl = reinterpret_cast<const float4 *>(p.locations) + row0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = tid + c * T;
            loc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            cf[c] = 0.5f;
            if (j < P) {
                loc[c] = ld_stream_f4(gl + j);
                cf[c] = ld_stream_f(p.confidences + row0 + j);
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = tid + c * T;
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
                    if (p.conf_out) p.conf_out[row0 + j] = cf[c];
                }
                s.row4col[j] = -1;
            }
            const float ce = boundary ? cf[c] : __fadd_rn(cf[c], kEps32);   // loss.py:74
            lc[c] = nplogf(ce);                                              // loss.py:21
            float w = __fsub_rn(1.0f, ce);                                   // loss.py:22-24
            if (w > 1.0f) w = 1.0f;
            if (w <= 0.0f) w = kEps32;
            l1[c] = nplogf(w);                                               // loss.py:25
            v[c] = 0.0;
            spc[c] = INF;
            r4c[c] = -1;
            pm[c] = 0;
        }
        const float4 *gg = reinterpret_cast<const float4 *>(p.gt) + static_cast<size_t>(b) * M;
        for (int i = tid; i < n; i += T) {
            s.gt[i] = gg[i];
            s.u[i] = 0.0;
            s.col4row[i] = -1;
        }
        block_sync<NWARPS>();

        // ---- one shortest augmenting path per GT row (rows = GT, columns = priors)
        bool failed = false;
        for (int cur = 0; cur < n && !failed; ++cur) {
            int i = cur, R = 0;
            double min_val = 0.0, ui = 0.0;
            unsigned scmask = invalid_mask;   // columns already scanned (or non-existent)
            for (;;) {
                const float4 g = s.gt[i];
                double best = INF;
                unsigned bpay = kPayNone;
                bool tie = false;
                if (R == 0) {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        if ((scmask >> c) & 1u) continue;
                        const float c32 = cost32(loc[c], g, half_alpha, lc[c], l1[c]);
                        if (!(c32 > -CUDART_INF_F)) status |= MBX_STATUS_INVALID_COST;   // NaN or -inf
                        const double r = __dsub_rn(static_cast<double>(c32), v[c]);     // (0 + C) - 0 - v
                        const double sp = r < INF ? r : INF;
                        spc[c] = sp;
                        pm[c] = 0;
                        const unsigned pay = (static_cast<unsigned>(tid + c * T) << 16) | static_cast<unsigned>(r4c[c] + 1);
                        if (sp < best) {
                            best = sp;
                            bpay = pay;
                            tie = false;
                        } else if (sp == best) {
                            tie = true;
                        }
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        if ((scmask >> c) & 1u) continue;
                        const float c32 = cost32(loc[c], g, half_alpha, lc[c], l1[c]);
                        const double r =
                            __dsub_rn(__dsub_rn(__dadd_rn(min_val, static_cast<double>(c32)), ui), v[c]);
                        if (r < spc[c]) {
                            spc[c] = r;
                            pm[c] = R;
                        }
                        const double sp = spc[c];
                        const unsigned pay = (static_cast<unsigned>(tid + c * T) << 16) | static_cast<unsigned>(r4c[c] + 1);
                        if (sp < best) {
                            best = sp;
                            bpay = pay;
                            tie = false;
                        } else if (sp == best) {
                            tie = true;
                        }
                    }
                }
                // ---- block-wide arg-min of (path cost, column); exact ties flagged
                const unsigned long long key = (bpay == kPayNone) ? ~0ull : ord64(best);
                unsigned hi = static_cast<unsigned>(key >> 32), lo = static_cast<unsigned>(key);
                unsigned pay = bpay;
                warp_argmin(hi, lo, pay, tie);
                if (NWARPS > 1) {
                    if (lane == 0)
                        s.part[pbuf * NWARPS + warp] =
                            make_int4(static_cast<int>(hi), static_cast<int>(lo), static_cast<int>(pay | (tie ? kPayTie : 0u)), 0);
                    __syncthreads();
                    int4 e = make_int4(-1, -1, -1, 0);
                    if (lane < NWARPS) e = s.part[pbuf * NWARPS + lane];
                    hi = static_cast<unsigned>(e.x);
                    lo = static_cast<unsigned>(e.y);
                    pay = static_cast<unsigned>(e.z);
                    tie = (pay != kPayNone) && (pay & kPayTie);
                    if (pay != kPayNone) pay &= ~kPayTie;
                    warp_argmin(hi, lo, pay, tie);
                    pbuf ^= 1;
                }
                const unsigned long long mkey = (static_cast<unsigned long long>(hi) << 32) | lo;
                if (pay == kPayNone || mkey >= ord64(INF)) {   // infeasible (scipy raises ValueError)
                    status |= MBX_STATUS_INFEASIBLE;
                    failed = true;
                    break;
                }
                min_val = unord64(mkey);
                int jstar = static_cast<int>(pay >> 16);
                int r4c_star = static_cast<int>(pay & 0x7fffu) - 1;
                if (tie) {
                    // scipy's rule among the columns AT the minimum: the LAST unassigned one in
                    // `remaining` order wins, else the FIRST assigned one (rare path).
                    unsigned long long k = ~0ull;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        if (((scmask >> c) & 1u) || !(spc[c] == min_val)) continue;
                        const int j = tid + c * T;
                        const int pos = replay_pos(j, R, P, s.rm_idx);
                        const unsigned k2 = (r4c[c] < 0) ? static_cast<unsigned>(P - 1 - pos) : static_cast<unsigned>(P + pos);
                        const unsigned long long kk = (static_cast<unsigned long long>(k2) << 32) |
                                                      (static_cast<unsigned>(j) << 16) | static_cast<unsigned>(r4c[c] + 1);
                        k = kk < k ? kk : k;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const unsigned long long t = __shfl_xor_sync(0xffffffffu, k, o);
                        k = t < k ? t : k;
                    }
                    if (NWARPS > 1) {
                        if (lane == 0) s.pk[warp] = k;
                        __syncthreads();
                        k = s.pk[0];
#pragma unroll
                        for (int w = 1; w < NWARPS; ++w) k = s.pk[w] < k ? s.pk[w] : k;
                        __syncthreads();
                    }
                    jstar = static_cast<int>((k >> 16) & 0xffffu);
                    r4c_star = static_cast<int>(k & 0x7fffu) - 1;
                }
                // ---- remove jstar from the scan set; its owner logs it
                const int cstar = jstar / T;
                if (jstar - cstar * T == tid) {
                    int pmv = 0;
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        if (c == cstar) pmv = pm[c];
                    scmask |= 1u << cstar;
                    s.rm_col[R] = jstar;
                    s.rm_idx[R] = replay_pos(jstar, R, P, s.rm_idx);
                    s.rm_pm[R] = pmv;
                    s.visit[R + 1] = r4c_star;
                }
                ++R;
                if (r4c_star < 0) break;   // jstar is the sink
                i = r4c_star;
                ui = s.u[i];
                if (NWARPS == 1) __syncwarp();
            }
            block_sync<NWARPS>();
            if (failed) break;
            // ---- dual update: v (owner registers), u (shared)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                if (((scmask & ~invalid_mask) >> c) & 1u) {
                    const double delta = __dsub_rn(min_val, spc[c]);
                    v[c] = __dsub_rn(v[c], delta);
                    // every scanned column except the sink is assigned, and its row was visited
                    if (r4c[c] >= 0) s.u[r4c[c]] = __dadd_rn(s.u[r4c[c]], delta);
                }
                spc[c] = INF;
            }
            if (tid == 0) {
                s.u[cur] = __dadd_rn(s.u[cur], min_val);
                // ---- augment along the path, from the sink back to row `cur`
                int k = R - 1;
                for (;;) {
                    const int m = s.rm_pm[k];
                    const int row = (m == 0) ? cur : s.visit[m];
                    const int col = s.rm_col[k];
                    s.row4col[col] = static_cast<short>(row);
                    s.col4row[row] = col;
                    if (m == 0) break;
                    k = m - 1;
                }
            }
            block_sync<NWARPS>();
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (!((invalid_mask >> c) & 1u)) r4c[c] = s.row4col[tid + c * T];
        }

        // ---- epilogue: mask, matched GT index, loss terms, gradients
        double acc_sq = 0.0, acc_conf = 0.0;
        int n_match = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = tid + c * T;
            if (j >= P) continue;
            const int r = r4c[c];
            if (p.mask) p.mask[row0 + j] = r >= 0 ? 1 : 0;
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
            if (p.d_loc) st_stream_f4(reinterpret_cast<float4 *>(p.d_loc) + row0 + j, dl);
            if (p.d_conf) p.d_conf[row0 + j] = dc;
        }
        if (p.stacked && !failed) {
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
            p.partials[2 * b] = a;
            p.partials[2 * b + 1] = cc;
            p.img_matched[b] = m;
        }
        block_sync<NWARPS>();   // shared state is reused by the next image
    }

    if (status) atomicOr(p.status, status);

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
    if (tid == 0) {
        double A = 0.0, Cc = 0.0, Mt = 0.0;
        for (int w = 0; w < NWARPS; ++w) {
            A += s.red[w];
            Cc += s.red[NWARPS + w];
            Mt += s.red[2 * NWARPS + w];
        }
        const double loc_loss = static_cast<double>(p.alpha) * (A / 2.0);   // loss.py:100
        const unsigned st = atomicOr(p.status, 0u);
        p.results[0] = static_cast<float>(loc_loss);
        p.results[1] = static_cast<float>(Cc);
        p.results[2] = static_cast<float>(st);
        p.results[3] = static_cast<float>(Mt);
        reinterpret_cast<double *>(p.results)[2] = loc_loss;
        reinterpret_cast<double *>(p.results)[3] = Cc;
        *p.ticket = 0u;    // workspace reusable by the next launch
        *p.status = 0u;
    }
}

namespace {

struct KernelInfo {
    size_t configured_smem = 0;
    int occ = 0;
    size_t occ_smem = 0;
};

template <int NWARPS, int C>
int launch_one(const MatchParams &p, cudaStream_t st) {
    static thread_local KernelInfo info;
    auto kern = mbx_match_loss_reg_kernel<NWARPS, C>;
    const size_t smem = rcarve(nullptr, nullptr, p.P, p.M, NWARPS, !(p.flags & MBX_FLAG_BOUNDARY));
    if (smem > static_cast<size_t>(max_smem_optin())) return MBX_E_TOO_LARGE;
    if (smem > info.configured_smem) {
        if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    static_cast<int>(smem)),
                               "cudaFuncSetAttribute(match_reg)"))
            return e;
        info.configured_smem = smem;
        info.occ = 0;
    }
    if (info.occ == 0 || info.occ_smem != smem) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&info.occ, kern, NWARPS * 32, smem);
        if (info.occ < 1) info.occ = 1;
        info.occ_smem = smem;
    }
    int grid = sm_count() * info.occ;
    if (grid > p.B) grid = p.B;
    kern<<<grid, NWARPS * 32, smem, st>>>(p);
    return check_cuda(cudaGetLastError(), "launch mbx_match_loss_reg_kernel");
}

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

}  // namespace

int launch_match_reg(const MatchParams &p, int force_warps, int force_cols, cudaStream_t st) {
    if (p.P > 65535 || p.M > 32766) return MBX_E_TOO_LARGE;
    int nwarps = force_warps;
    if (nwarps == 0) {
        // latency mode (few images: every CTA has an SM to itself) uses wide CTAs; throughput
        // mode keeps CTAs narrow so several images share an SM.
        const bool latency = p.B <= 2 * sm_count();
        if (latency)
            nwarps = p.P <= 128 ? 2 : (p.P <= 320 ? 4 : (p.P <= 1024 ? 8 : 16));
        else
            nwarps = p.P <= 96 ? 1 : (p.P <= 384 ? 2 : (p.P <= 768 ? 4 : 8));
    }
    int cols = force_cols ? force_cols : (p.P + nwarps * 32 - 1) / (nwarps * 32);
    while (!force_warps && cols > 8 && nwarps < 16) {
        nwarps *= 2;
        cols = (p.P + nwarps * 32 - 1) / (nwarps * 32);
    }
    if (cols * nwarps * 32 < p.P) return MBX_E_TOO_LARGE;
    switch (nwarps) {
        case 1: return launch_cols<1>(p, cols, st);
        case 2: return launch_cols<2>(p, cols, st);
        case 4: return launch_cols<4>(p, cols, st);
        case 8: return launch_cols<8>(p, cols, st);
        case 16: return launch_cols<16>(p, cols, st);
        default: return MBX_E_TOO_LARGE;
    }
}

}  // namespace mbx
