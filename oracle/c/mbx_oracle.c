/* TEST INFRASTRUCTURE -- plain-C CPU restatement of the Multibox matching path.
 *
 * Checker only: linked/loaded by tests/, __graft_entry__.smoke() and bench.py's
 * CPU-baseline legs.  The product (multibox_b200/) never loads this library.
 *
 * Restates
 *   - the cost matrix + assignment loop of the reference's compute_assignments
 *     (reference loss.py:8-53),
 *   - the third-party arithmetic that loop calls and that is not in the
 *     reference checkout:
 *       * scipy.optimize.linear_sum_assignment (reference loss.py:2,40; pinned
 *         scipy==0.17.0 in requirements.txt:5, this image ships 1.18.1): the
 *         published algorithm of scipy >= 1.4 -- D. F. Crouse, "On implementing
 *         2D rectangular assignment algorithms", IEEE TAES 52(4), 2016
 *         (shortest augmenting paths, dual variables u/v, one augmentation per
 *         row, tall matrices transposed first) -- including scipy's scan order
 *         and tie rule (columns scanned from a 'remaining' list filled in
 *         reverse; among equal reduced costs an unassigned column wins).
 *         Pinned against the scipy in this image on tie-heavy integer matrices
 *         and random float matrices (tests/test_oracle_c.py).
 *       * numpy's float32 np.log (reference loss.py:21,25): numpy's SIMD
 *         kernel is a degree-5/5 rational minimax approximation evaluated with
 *         FMAs (not correctly rounded: up to ~3.8 ulp).  orc_nplogf restates it
 *         and is bit-equal to np.log on every positive finite float32 on this
 *         image's AVX-512 host (tests/test_oracle_c.py samples the range;
 *         oracle/gen_golden.py --exhaustive walks all 2^31 values).
 *       * numpy's np.linalg.norm(x, axis=1) on [P,4] float32: sqrt of the
 *         left-to-right fp32 sum of squares.
 *
 * Build: see oracle/Makefile (gcc -O2 -mfma -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_INVALID 1     /* NaN or -inf cost entry (scipy: ValueError) */
#define ORC_INFEASIBLE 2  /* no finite assignment (scipy: ValueError)   */

/* ---- numpy float32 log ------------------------------------------------- */
float orc_nplogf(float x_in)
{
    static const float P0 = 0.000000000000000000000e+00f, P1 = 9.999999999999998702752e-01f,
                       P2 = 2.112677543073053063722e+00f, P3 = 1.480000633576506585156e+00f,
                       P4 = 3.808837741388407920751e-01f, P5 = 2.589979117907922693523e-02f;
    static const float Q0 = 1.000000000000000000000e+00f, Q1 = 2.612677543073109236779e+00f,
                       Q2 = 2.453006071784736363091e+00f, Q3 = 9.864942958519418960339e-01f,
                       Q4 = 1.546476374983906719538e-01f, Q5 = 5.875095403124574342950e-03f;
    if (x_in != x_in) return x_in;
    if (x_in < 0.0f) return NAN;
    if (x_in == 0.0f) return -INFINITY;
    if (isinf(x_in)) return x_in;
    int e;
    float m = frexpf(x_in, &e);               /* m in [0.5, 1) */
    float ef = (float)e;
    if (m <= 0.70710678118654752440f) { m = m + m; ef -= 1.0f; }
    float x = m - 1.0f;
    float n = fmaf(P5, x, P4); n = fmaf(n, x, P3); n = fmaf(n, x, P2); n = fmaf(n, x, P1); n = fmaf(n, x, P0);
    float d = fmaf(Q5, x, Q4); d = fmaf(d, x, Q3); d = fmaf(d, x, Q2); d = fmaf(d, x, Q1); d = fmaf(d, x, Q0);
    float p = n / d;
    return fmaf(ef, 0.693147180559945309417232121458176568f, p);
}

void orc_nplogf_array(const float *in, float *out, int64_t n)
{
    for (int64_t i = 0; i < n; i++) out[i] = orc_nplogf(in[i]);
}

/* ---- rectangular LSAP (Crouse / scipy) ---------------------------------- */
/* instrumentation: Dijkstra steps (full column scans) and column visits since the last reset */
static int64_t g_scan_steps = 0, g_col_visits = 0;
void orc_counters(int64_t *steps, int64_t *visits, int reset)
{
    if (steps) *steps = g_scan_steps;
    if (visits) *visits = g_col_visits;
    if (reset) { g_scan_steps = 0; g_col_visits = 0; }
}

/* cost is row-major [nr, nc] with nr <= nc (caller transposes tall inputs).
 * col4row[nr] receives the column assigned to each row. */
static int lsap_wide(int64_t nr, int64_t nc, const double *cost, int64_t *col4row)
{
    double *u = calloc(nr, sizeof(double)), *v = calloc(nc, sizeof(double));
    double *spc = malloc(nc * sizeof(double));
    int64_t *path = malloc(nc * sizeof(int64_t)), *row4col = malloc(nc * sizeof(int64_t));
    int64_t *remaining = malloc(nc * sizeof(int64_t));
    char *SR = malloc(nr), *SC = malloc(nc);
    int rc = ORC_OK;
    for (int64_t j = 0; j < nc; j++) { path[j] = -1; row4col[j] = -1; }
    for (int64_t i = 0; i < nr; i++) col4row[i] = -1;

    for (int64_t cur = 0; cur < nr && rc == ORC_OK; cur++) {
        /* shortest augmenting path from row `cur` */
        double min_val = 0.0;
        int64_t num_remaining = nc, i = cur, sink = -1;
        for (int64_t it = 0; it < nc; it++) remaining[it] = nc - it - 1;   /* reverse fill */
        memset(SR, 0, nr); memset(SC, 0, nc);
        for (int64_t j = 0; j < nc; j++) spc[j] = INFINITY;
        while (sink == -1) {
            int64_t index = -1;
            double lowest = INFINITY;
            SR[i] = 1;
            g_scan_steps++; g_col_visits += num_remaining;
            for (int64_t it = 0; it < num_remaining; it++) {
                int64_t j = remaining[it];
                double r = min_val + cost[i * nc + j] - u[i] - v[j];
                if (r < spc[j]) { path[j] = i; spc[j] = r; }
                if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) {
                    lowest = spc[j]; index = it;
                }
            }
            min_val = lowest;
            if (min_val == INFINITY) { rc = ORC_INFEASIBLE; break; }
            int64_t j = remaining[index];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            SC[j] = 1;
            remaining[index] = remaining[--num_remaining];
        }
        if (rc != ORC_OK) break;
        /* dual update */
        u[cur] += min_val;
        for (int64_t r = 0; r < nr; r++)
            if (SR[r] && r != cur) u[r] += min_val - spc[col4row[r]];
        for (int64_t j = 0; j < nc; j++)
            if (SC[j]) v[j] -= min_val - spc[j];
        /* augment */
        int64_t j = sink;
        for (;;) {
            int64_t r = path[j];
            row4col[j] = r;
            int64_t t = col4row[r]; col4row[r] = j; j = t;
            if (r == cur) break;
        }
    }
    free(u); free(v); free(spc); free(path); free(row4col); free(remaining); free(SR); free(SC);
    return rc;
}

/* General entry with scipy's conventions: cost [nr, nc] row-major; outputs
 * row_ind/col_ind of length min(nr, nc), row_ind ascending. */
int orc_lsap(int64_t nr, int64_t nc, const double *cost, int64_t *row_ind, int64_t *col_ind)
{
    if (nr == 0 || nc == 0) return ORC_OK;
    for (int64_t k = 0; k < nr * nc; k++)
        if (cost[k] != cost[k] || cost[k] == -INFINITY) return ORC_INVALID;
    if (nc >= nr) {
        int rc = lsap_wide(nr, nc, cost, col_ind);
        for (int64_t i = 0; i < nr; i++) row_ind[i] = i;
        return rc;
    }
    /* tall: solve the transpose, then order by the original row index */
    double *t = malloc((size_t)(nr * nc) * sizeof(double) + 8);
    for (int64_t i = 0; i < nr; i++)
        for (int64_t j = 0; j < nc; j++) t[j * nr + i] = cost[i * nc + j];
    int64_t *c4r = malloc(nc * sizeof(int64_t));
    int rc = lsap_wide(nc, nr, t, c4r);
    if (rc == ORC_OK) {
        /* argsort of c4r (distinct values): counting placement */
        int64_t *owner = malloc(nr * sizeof(int64_t));
        for (int64_t i = 0; i < nr; i++) owner[i] = -1;
        for (int64_t j = 0; j < nc; j++) owner[c4r[j]] = j;
        int64_t k = 0;
        for (int64_t i = 0; i < nr; i++)
            if (owner[i] >= 0) { row_ind[k] = i; col_ind[k] = owner[i]; k++; }
        free(owner);
    }
    free(t); free(c4r);
    return rc;
}

/* ---- reference loss.py:21-25 and :33-35 ----------------------------------- */
void orc_log_terms(const float *conf, float *log_c, float *log_1mc, int64_t n)
{
    for (int64_t i = 0; i < n; i++) {
        log_c[i] = orc_nplogf(conf[i]);
        float v = 1.0f - conf[i];
        if (v > 1.0f) v = 1.0f;
        if (v <= 0.0f) v = 1e-10f;            /* float32(SMALL_EPSILON) */
        log_1mc[i] = orc_nplogf(v);
    }
}

/* fp32 cost of (prior p, gt g) widened to double, exact numpy op order:
 * (alpha/2) * (sqrt(((d0^2+d1^2)+d2^2)+d3^2))**2 - log_c + log_1mc */
static inline double cost_entry(const float *loc, const float *g, float half_alpha, float lc, float l1)
{
    float d0 = loc[0] - g[0], d1 = loc[1] - g[1], d2 = loc[2] - g[2], d3 = loc[3] - g[3];
    float s = d0 * d0; s = s + d1 * d1; s = s + d2 * d2; s = s + d3 * d3;
    float nrm = sqrtf(s);
    float c = half_alpha * (nrm * nrm);
    c = c - lc;
    c = c + l1;
    return (double)c;
}

/* Whole-batch compute_assignments.  loc [B*P,4] (prior added), conf [B*P]
 * (epsilon added), gt [B,M,4], num_gt [B].  Outputs: mask [B*P] (0/1),
 * gt_idx [B*P] (-1 or gt index), stacked [sum n,4] in (image, prior) order,
 * *n_stacked.  Returns ORC_*. */
int orc_compute_assignments(const float *loc, const float *conf, const float *gt, const int32_t *num_gt,
                            int64_t B, int64_t P, int64_t M, float alpha,
                            int32_t *mask, int32_t *gt_idx, float *stacked, int64_t *n_stacked)
{
    float half_alpha = alpha / 2.0f;
    float *lc = malloc(B * P * sizeof(float)), *l1 = malloc(B * P * sizeof(float));
    double *C = malloc((size_t)P * (M > 0 ? M : 1) * sizeof(double));
    int64_t *ri = malloc((M + 1) * sizeof(int64_t)), *ci = malloc((M + 1) * sizeof(int64_t));
    int rc = ORC_OK;
    int64_t ns = 0;
    orc_log_terms(conf, lc, l1, B * P);
    for (int64_t k = 0; k < B * P; k++) { mask[k] = 0; gt_idx[k] = -1; }
    for (int64_t b = 0; b < B && rc == ORC_OK; b++) {
        int64_t n = num_gt[b];
        const float *L = loc + b * P * 4, *G = gt + b * M * 4;
        for (int64_t p = 0; p < P; p++)
            for (int64_t j = 0; j < n; j++)
                C[p * n + j] = cost_entry(L + 4 * p, G + 4 * j, half_alpha, lc[b * P + p], l1[b * P + p]);
        rc = orc_lsap(P, n, C, ri, ci);
        if (rc != ORC_OK) break;
        int64_t cnt = P < n ? P : n;
        for (int64_t k = 0; k < cnt; k++) {
            mask[b * P + ri[k]] = 1;
            gt_idx[b * P + ri[k]] = (int32_t)ci[k];
            memcpy(stacked + 4 * ns, G + 4 * ci[k], 4 * sizeof(float));
            ns++;
        }
    }
    *n_stacked = ns;
    free(lc); free(l1); free(C); free(ri); free(ci);
    return rc;
}

/* The cost matrix alone, for bit-level comparison with numpy: C [P, n] doubles. */
void orc_cost_matrix(const float *loc, const float *conf, const float *gt, int64_t P, int64_t n,
                     float alpha, double *C)
{
    float half_alpha = alpha / 2.0f;
    float *lc = malloc(P * sizeof(float)), *l1 = malloc(P * sizeof(float));
    orc_log_terms(conf, lc, l1, P);
    for (int64_t p = 0; p < P; p++)
        for (int64_t j = 0; j < n; j++)
            C[p * n + j] = cost_entry(loc + 4 * p, gt + 4 * j, half_alpha, lc[p], l1[p]);
    free(lc); free(l1);
}

/* ---- numeric check of the product's cheap cost bound -------------------------
 * The CUDA kernels skip the exact cost wherever a 4-FMA approximation a(i,j) + G_i
 * minus a proven margin already exceeds what is needed (multibox_b200/csrc/mbx_bound.h
 * holds the formulas and the error analysis; the SAME header is compiled here as plain
 * C).  orc_bound_max_ratio returns max over all (prior, gt) pairs of
 *     |c_exact - (a + G)| / (m_prior + m_gt)        (must stay <= 1; evaluated in double)
 * with c_exact from cost_entry() above; pairs whose margin is +inf (pruning disabled) are
 * skipped and counted in *n_unbounded.  A NaN ratio (finite margin, non-finite values)
 * is reported as +inf. */
#include "../../multibox_b200/csrc/mbx_bound.h"

/* lc_w: the (possibly approximate) log(c) the cheap form is built with -- the kernels use a fast
 * hardware log there; lc: the exact numpy log the cost entry uses. */
double orc_bound_max_ratio(const float *loc, const float *lc, const float *lc_w, const float *l1, const float *gt,
                           int64_t P, int64_t n, float alpha, int64_t *n_unbounded)
{
    const float h = alpha / 2.0f;
    double worst = 0.0;
    int64_t unb = 0;
    for (int64_t i = 0; i < n; i++) {
        const float *g = gt + 4 * i;
        float gp[4], G, mg;
        mbx_bound_row(g[0], g[1], g[2], g[3], h, gp, &G, &mg);
        for (int64_t p = 0; p < P; p++) {
            const float *l = loc + 4 * p;
            const float L = fmaxf(fmaxf(fabsf(l[0]), fabsf(l[1])), fmaxf(fabsf(l[2]), fabsf(l[3])));
            float Lx = L;
            for (int k = 0; k < 4; k++)
                if (!(fabsf(l[k]) <= 3.4028234663852886e38f)) Lx = NAN;
            const float T = fabsf(lc_w[p]) + fabsf(l1[p]);
            const float m = mbx_bound_margin_col(Lx, T, h);
            if (isinf(m) || isinf(mg)) { unb++; continue; }
            const float w = mbx_bound_w(l[0], l[1], l[2], l[3], h, lc_w[p], l1[p]);
            const float a = mbx_bound_a(l[0], l[1], l[2], l[3], gp[0], gp[1], gp[2], gp[3], w);
            const double c = cost_entry(l, g, h, lc[p], l1[p]);
            const double err = fabs(c - ((double)a + (double)G));
            const double ratio = err / ((double)m + (double)mg);
            if (!(ratio <= worst)) worst = (ratio != ratio) ? INFINITY : ratio;
        }
    }
    if (n_unbounded) *n_unbounded = unb;
    return worst;
}
