// Softmax arithmetic of the attention kernel (attention6.cu).
//
// The softmax warps of those kernels are instruction-issue bound (ncu, profiles/r01/README.md: no pipe above 45 %, ~40 %
// of the samples are dependency waits), so the per-element instruction count is what matters:
//   scale+shift  fma.rn.f32x2   1 instruction per PAIR of scores (FFMA2, Blackwell packed fp32)
//   exp2         ex2.approx     1 per element (MUFU; at these rates the XU pipe stays below 50 %)
//   row sum      add.rn.f32x2   1 per pair (FADD2)
//   bf16 pack    cvt.rn.bf16x2  1 per pair
//   row max      3-input max    1 per pair (FMNMX3, formed by ptxas)
// = 3 instructions per score. POLY > 0 moves the second exp2 of every POLY-th pair to the FMA pipe (Cody-Waite split +
// cubic, ~9 instructions) to relieve the MUFU pipe (16 ex2 / clk / SM) once the kernel runs fast enough to load it.
#pragma once
#include <cstdint>

namespace mmdp {

__device__ __forceinline__ uint64_t f32x2_pack(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f32x2_unpack(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f32x2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t f32x2_add(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float ex2_mufu(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_poly(float x) {  // 2^x on the FMA pipe: Cody-Waite split (magic-number rounding) + cubic on [-0.5, 0.5]
    x = fmaxf(x, -126.0f);
    const float t = x + 12582912.0f;
    const float f = x - (t - 12582912.0f);
    float p = fmaf(0.0551716573536396f, f, 0.2426111251115799f);
    p = fmaf(p, f, 0.6932609677314758f);
    p = fmaf(p, f, 0.9999280571937561f);
    return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}

// P = 2^(s * c + mneg) for N (even) scores held as raw fp32 bits in sv; writes N/2 packed bf16 pairs to pk and adds the
// fp32 row sum into four packed accumulators (8 independent chains).
template <int N, int POLY = 0>
__device__ __forceinline__ void softmax_exp_block(const uint32_t* sv, float scale_log2, float mneg, uint32_t* pk, uint64_t (&acc)[4]) {
    const uint64_t c2 = f32x2_pack(scale_log2, scale_log2), m2 = f32x2_pack(mneg, mneg);
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        float x0, x1;
        f32x2_unpack(f32x2_fma(f32x2_pack(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), c2, m2), x0, x1);
        const float p0 = ex2_mufu(x0);
        const float p1 = (POLY > 0 && (i % (POLY > 0 ? POLY : 1)) == 0) ? ex2_poly(x1) : ex2_mufu(x1);
        acc[i & 3] = f32x2_add(acc[i & 3], f32x2_pack(p0, p1));
        pk[i] = pack_bf16x2(p0, p1);
    }
}
__device__ __forceinline__ float f32x2_sum4(const uint64_t (&acc)[4]) {
    float a0, a1, b0, b1;
    f32x2_unpack(f32x2_add(acc[0], acc[1]), a0, a1);
    f32x2_unpack(f32x2_add(acc[2], acc[3]), b0, b1);
    return (a0 + a1) + (b0 + b1);
}

}  // namespace mmdp
