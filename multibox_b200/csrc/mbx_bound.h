// multibox_b200 -- the cheap bound of the matching cost (shared by the CUDA kernels and, as plain C,
// by the CPU test that checks the error analysis numerically: tests/test_cost_bound.py).
//
// Exact cost (reference loss.py:35, numpy operation order, fp32; cost32() in mbx_match.cuh):
//     c(i,j) = fl(fl(h * fl(nrm^2)) - lc_j) + l1_j),  nrm = fl(sqrt(fl-sum of fl((l_jk - g_ik)^2)))
// with h = alpha/2, l_j the predicted box of prior j, g_i the GT box i, lc_j = log(c_j), l1_j = log(1-c_j).
//
// Cheap form (4 FMAs per (GT, prior) pair): expand the square,
//     c*(i,j) = [h*|l_j|^2 + (l1_j - lc_j)] + sum_k l_jk * (-2h g_ik) + h*|g_i|^2
//             =            w_j              +        l_j . gp_i       +    G_i
//     a(i,j)  = fma(l_j0, gp_i0, fma(l_j1, gp_i1, fma(l_j2, gp_i2, fma(l_j3, gp_i3, w_j))))
// w_j is a per-prior constant, gp_i / G_i per-GT constants, all fp32.
//
// Error bound (u = 2^-24, standard model fl(x) = x(1+d), |d| <= u; L_j = max_k |l_jk|, Gm_i = max_k |g_ik|,
// T_j = |lc_j| + |l1_j|, S* = sum_k (l_jk - g_ik)^2 <= 4 (L_j + Gm_i)^2):
//   exact path : each squared difference carries 3 roundings, the sequential sum 3 more, sqrt 1, the re-squaring
//                1, the product with h 1 -> |A - h S*| <= g10 h S*; the two additions give
//                |c - c*| <= u (12.1 h S* + 2.1 |lc| + 1.1 |l1|)               <= u (48.4 h (L+Gm)^2 + 2.1 T)
//   cheap path : |w - w*| <= u (5.1 h |l|^2 + 2.1 T); gp carries 1 rounding, the FMA chain 4:
//                |a - (c* - G*)| <= u (40.1 h L Gm + 36.8 h L^2 + 6.2 T);  |G - G*| <= u 20.1 h Gm^2
//   together   : |c - (a + G)| <= u (115.4 h (L+Gm)^2 + 8.3 T) <= u (231 h L^2 + 231 h Gm^2 + 8.3 T)
// The cheap form may be fed an APPROXIMATE log(c): lc' with |lc' - lc| <= 2^-21 + 2^-19 |lc'| (the kernel uses
// __logf: absolute error <= 2^-21.41 on [0.5, 2], <= 3 ulp elsewhere -- CUDA C++ Programming Guide, checked
// on every float32 by mbx_debug_fastlog_violations -- against numpy's log, itself within 4 ulp of the true
// one); w_j then moves by at most that much, which the T term and the constant below absorb
// (8.3 u T + 2^-19 T + 2^-21 <= 2^-18 T + 2^-20).
// The margins below use 256 / 256 (>= 10 % head room, which also covers the round-to-nearest
// arithmetic of the margins themselves) plus 2^-100 absolute for subnormal intermediates:
//     |c(i,j) - (a(i,j) + G_i)| <= m_j + mg_i,   m_j = 2^-16 |h| L_j^2 + 2^-18 T_j + 2^-20,  mg_i = 2^-16 |h| Gm_i^2
// with T_j = |lc'_j| + |l1_j| taken from the values the cheap form was built with.
// A margin above 2^60 or not finite (non-finite inputs, overflow) becomes +inf: nothing is pruned with it.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define MBX_BOUND_FN __host__ __device__ __forceinline__
#else
#define MBX_BOUND_FN static inline
#endif

#define MBX_BOUND_CAP 1.152921504606847e18f /* 2^60 */

MBX_BOUND_FN float mbx_bound_cap(float m) { return (m < MBX_BOUND_CAP) ? m : INFINITY; }

// per-prior constant w_j = fl(h*|l|^2 + fl(l1 - lc))
MBX_BOUND_FN float mbx_bound_w(float l0, float l1_, float l2, float l3, float h, float lc, float l1) {
    float q = l0 * l0;
    q = fmaf(l1_, l1_, q);
    q = fmaf(l2, l2, q);
    q = fmaf(l3, l3, q);
    return fmaf(h, q, l1 - lc);
}

// per-prior margin from L = max_k |l_k| and T = |lc| + |l1| (or any upper bounds of them)
MBX_BOUND_FN float mbx_bound_margin_col(float L, float T, float h) {
    const float m = fmaf(1.52587890625e-05f * fabsf(h), L * L, fmaf(3.814697265625e-06f, T, 9.5367431640625e-07f));
    return mbx_bound_cap(m);
}

// per-GT constants: gp = -2h g (4 values), G = h*|g|^2, mg = 2^-16 |h| max|g|^2
MBX_BOUND_FN void mbx_bound_row(float g0, float g1, float g2, float g3, float h, float *gp, float *G, float *mg) {
    const float m2h = -2.0f * h;
    gp[0] = m2h * g0;
    gp[1] = m2h * g1;
    gp[2] = m2h * g2;
    gp[3] = m2h * g3;
    float q = g0 * g0;
    q = fmaf(g1, g1, q);
    q = fmaf(g2, g2, q);
    q = fmaf(g3, g3, q);
    *G = h * q;
    const float gm = fmaxf(fmaxf(fabsf(g0), fabsf(g1)), fmaxf(fabsf(g2), fabsf(g3)));
    float m = (1.52587890625e-05f * fabsf(h)) * (gm * gm);
    // (fmaxf drops NaNs: a NaN / inf coordinate must still disable pruning for this row)
    if (!(fabsf(g0) <= 3.4028234663852886e38f) || !(fabsf(g1) <= 3.4028234663852886e38f) ||
        !(fabsf(g2) <= 3.4028234663852886e38f) || !(fabsf(g3) <= 3.4028234663852886e38f))
        m = INFINITY;
    *mg = mbx_bound_cap(m);
}

// the cheap form: a(i,j) + G_i approximates c(i,j) within m_j + mg_i
MBX_BOUND_FN float mbx_bound_a(float l0, float l1, float l2, float l3, float gp0, float gp1, float gp2, float gp3,
                               float w) {
    return fmaf(l0, gp0, fmaf(l1, gp1, fmaf(l2, gp2, fmaf(l3, gp3, w))));
}
