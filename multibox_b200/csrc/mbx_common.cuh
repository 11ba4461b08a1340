// multibox_b200 -- shared device helpers (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/multibox_b200.h"

namespace mbx {

constexpr float kEps32 = 1e-10f;   // float32(SMALL_EPSILON), reference loss.py:6

void set_error(const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);
int sm_count();
int max_smem_optin();

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// 1-D TMA bulk copy global -> shared (SASS: UBLKCP); dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// streaming (read-once) vector loads / stores that stay out of L1
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream_f(const float *p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float4 *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// ---------------------------------------------------------------- numpy float32 log
// Bit-exact restatement of numpy's SIMD float32 log kernel (the np.log of
// reference loss.py:21,25 as it executes on AVX2/AVX-512 hosts): mantissa in
// [1/sqrt2, sqrt2), log(1+x) ~ P(x)/Q(x) (degree 5/5, Horner with FMAs), one
// IEEE division, + e*ln2 by FMA.  Pinned against np.log over every positive
// finite float32 by oracle/gen_golden.py --exhaustive (via the C twin in
// oracle/c/mbx_oracle.c) and against the CUDA build by tests/test_gpu_match.py.
__device__ __forceinline__ float nplogf(float x_in) {
    const float P0 = 0.000000000000000000000e+00f, P1 = 9.999999999999998702752e-01f,
                P2 = 2.112677543073053063722e+00f, P3 = 1.480000633576506585156e+00f,
                P4 = 3.808837741388407920751e-01f, P5 = 2.589979117907922693523e-02f;
    const float Q0 = 1.000000000000000000000e+00f, Q1 = 2.612677543073109236779e+00f,
                Q2 = 2.453006071784736363091e+00f, Q3 = 9.864942958519418960339e-01f,
                Q4 = 1.546476374983906719538e-01f, Q5 = 5.875095403124574342950e-03f;
    if (!(x_in > 0.0f) || x_in == __int_as_float(0x7f800000)) {
        if (x_in == 0.0f) return __int_as_float(0xff800000);   // -inf
        if (x_in < 0.0f) return __int_as_float(0x7fc00000);    // nan
        return x_in;                                            // +inf, nan
    }
    int e;
    float m;
    uint32_t bits = __float_as_uint(x_in);
    if (bits >= 0x00800000u) {   // normal: frexp by bit surgery, m in [0.5, 1)
        e = static_cast<int>(bits >> 23) - 126;
        m = __uint_as_float((bits & 0x007fffffu) | 0x3f000000u);
    } else {
        m = frexpf(x_in, &e);
    }
    float ef = static_cast<float>(e);
    if (m <= 0.70710678118654752440f) {
        m = __fadd_rn(m, m);
        ef = __fsub_rn(ef, 1.0f);
    }
    float x = __fsub_rn(m, 1.0f);
    float n = __fmaf_rn(P5, x, P4);
    n = __fmaf_rn(n, x, P3);
    n = __fmaf_rn(n, x, P2);
    n = __fmaf_rn(n, x, P1);
    n = __fmaf_rn(n, x, P0);
    float d = __fmaf_rn(Q5, x, Q4);
    d = __fmaf_rn(d, x, Q3);
    d = __fmaf_rn(d, x, Q2);
    d = __fmaf_rn(d, x, Q1);
    d = __fmaf_rn(d, x, Q0);
    float p = __fdiv_rn(n, d);
    return __fmaf_rn(ef, 0.693147180559945309417232121458176568f, p);
}

// Out-of-line copy for the rare call sites (exact log of a candidate's confidence, saturated
// confidences in the epilogue): keeps ~45 instructions per site out of the hot instruction stream.
static __device__ __noinline__ float nplogf_cold(float x) { return nplogf(x); }

// sigmoid of reference model.py:322 (fp32; tolerance-checked, not bit-pinned)
__device__ __forceinline__ float sigmoidf_(float z) { return __fdiv_rn(1.0f, 1.0f + expf(-z)); }

// ---------------------------------------------------------------- per-head input layout
// One detection head of the reference network (model.py:198-316): its conv outputs, NHWC, are
// [B, g, g, K*4] / [B, g, g, K] = [B, pri, 4] / [B, pri] with pri = g*g*K consecutive priors.
struct HeadTab {
    const float *loc, *conf;
    float *dloc, *dconf;
    int pri, off;   // priors of this head; index of its first prior in the concatenated order
};

// The head table staged in shared memory (dynamic indexing of kernel parameters would force a
// local-memory copy of the whole parameter block).  Call with all threads, then barrier.
template <typename Params>
__device__ __forceinline__ void stage_heads(const Params &p, HeadTab *sh) {
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < MBX_MAX_HEADS; ++k) sh[k] = p.heads[k];
    }
}
// element index of prior j of image b inside its head's tensors; h = the head
__device__ __forceinline__ size_t head_elem(const HeadTab *sh, int nheads, int j, int b, int &h) {
    h = 0;
    for (int k = 1; k < nheads; ++k) h += (j >= sh[k].off) ? 1 : 0;
    return static_cast<size_t>(b) * sh[h].pri + (j - sh[h].off);
}


// ---------------------------------------------------------------- warp / block reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace mbx
