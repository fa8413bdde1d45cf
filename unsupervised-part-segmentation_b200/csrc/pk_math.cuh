// Packed fp32x2 evaluation of the canonical exp / log / TPS radial basis (device only, sm_100a).
//
// Blackwell issues `fma.rn.f32x2` (SASS FFMA2): one instruction, two correctly rounded fp32
// results.  The canonical sequences of canon_math.cuh are chains of single-rounded mul / add,
// so evaluating TWO independent arguments side by side halves the instruction count of the
// issue-bound kernels (K1's 8 radial-basis logs per pixel, the K exps per pixel of the softmax
// kernels) while every lane result stays bit-identical to the scalar form and to oracle/canon.py.
//
// ptxas 12.9 contracts `mul.rn.f32x2` followed by `add.rn.f32x2` into one FFMA2 -- also for the
// __fmul2_rn/__fadd2_rn intrinsics, also with --fmad=false, and also when they are written as
// fma(a, b, -0.0) and fma(a, 1.0, c) with literal constants.  That would change the rounding.
// The single-rounded packed ops below are therefore FFMA2s whose neutral operand (1.0 / -0.0 /
// -1.0) comes from constant memory that ptxas cannot see through:
//     mul2(a,b) = fma(a, b, -0)   add2(a,b) = fma(a, 1, b)   sub2(a,b) = fma(b, -1, a)
// each of which is exactly the correctly rounded product / sum / difference (signed zeros
// included: x*y + (-0) keeps the sign of a zero product; a*1 + b is a + b).
#pragma once
#include "canon_math.cuh"

#if defined(__CUDACC__)
namespace ups {
namespace pk {

typedef unsigned long long f2;  // {lo, hi} fp32 pair in one 64-bit register pair

struct NeutralConsts { f2 one, nzero, none; };
// one definition per translation unit (static): no relocatable device code needed
static __constant__ NeutralConsts NEUTRAL = {0x3f8000003f800000ull, 0x8000000080000000ull, 0xbf800000bf800000ull};

__device__ __forceinline__ f2 pack(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f2 splat(float v) { return pack(v, v); }
__device__ __forceinline__ void unpack(f2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// A constant-bank value lives in a UNIFORM register, and an FFMA2 takes at most one uniform-or-
// immediate operand: fma(y, one, imm) or fma(T_uniform, r, nzero) then costs two MOVs to copy the
// neutral operand into a vector register pair -- at every add of the Horner chains.  ptxas proves
// uniformity through loads from uniform addresses too, so the three values are made formally
// lane-dependent: OR-ed with (%laneid + %smid) >> 16, which is 0 everywhere but not to the compiler
// (ptxas folds %laneid >> 5 by itself).
__device__ __forceinline__ unsigned lane_zero() {
    unsigned l, sm;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    return (l + sm) >> 16;
}
__device__ __forceinline__ f2 vector_resident(unsigned bits, unsigned z) {
    f2 r;
    const unsigned b = bits | z;
    asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "r"(b));
    return r;
}

// the three neutral operands, built once per kernel in vector registers
struct Ops {
    f2 one, nzero, none;
    __device__ __forceinline__ Ops() {
        const unsigned z = lane_zero();
        one = vector_resident(0x3f800000u, z);
        nzero = vector_resident(0x80000000u, z);
        none = vector_resident(0xbf800000u, z);
    }
    __device__ __forceinline__ f2 mul(f2 a, f2 b) const { return fma2(a, b, nzero); }
    __device__ __forceinline__ f2 add(f2 a, f2 b) const { return fma2(a, one, b); }
    __device__ __forceinline__ f2 sub(f2 a, f2 b) const { return fma2(b, none, a); }
    // fl(fl(a*b) + c): the canonical (uncontracted) multiply-add
    __device__ __forceinline__ f2 muladd(f2 a, f2 b, f2 c) const { return add(mul(a, b), c); }
};

// (float)e for |e| < 2^22 without the conversion pipe: bits(1.5*2^23 + e) = 0x4B400000 + e
__device__ __forceinline__ float small_int_as_magic(int e) { return __int_as_float(0x4B400000 + e); }
constexpr float MAGIC = 12582912.0f;

// log_canon of 2*NP arguments (normal positive), NP independent pairs advanced in lockstep: the
// Horner chain is one dependent FFMA2 after another, and ptxas does not interleave independent
// chains by itself, so the source order does it (NP chains in flight hide the FFMA2 latency).
// Lane results == canon_math.cuh::log_canon.
template <int NP>
__device__ __forceinline__ void log_canon2n(const Ops& o, const f2 (&x)[NP], f2 (&out)[NP]) {
    f2 m[NP], ef[NP], z[NP], y[NP];
#pragma unroll
    for (int n = 0; n < NP; ++n) {
        float x0, x1;
        unpack(x[n], x0, x1);
        const int b0 = __float_as_int(x0), b1 = __float_as_int(x1);
        // canonical split: mantissa m in [0.5, 1), e = E - 126; if m < sqrt(1/2): m <- 2m, e <- e - 1.
        // For a positive normal x that is one subtraction and one shift: with S = bits(sqrt(1/2)f) =
        // (126 << 23) + 0x3504F3, (b - S) >> 23 = E - 126 - (mantissa bits < 0x3504F3), and removing
        // that exponent from b leaves bits(m) or bits(2m).  Same e and t as the compare-and-select form.
        const int e0 = (b0 - 0x3F3504F3) >> 23, e1 = (b1 - 0x3F3504F3) >> 23;
        const float t0 = __int_as_float(b0 - (e0 << 23));
        const float t1 = __int_as_float(b1 - (e1 << 23));
        m[n] = o.sub(pack(t0, t1), o.one);
        ef[n] = o.sub(pack(small_int_as_magic(e0), small_int_as_magic(e1)), splat(MAGIC));  // exact
    }
#pragma unroll
    for (int n = 0; n < NP; ++n) z[n] = o.mul(m[n], m[n]);
#pragma unroll
    for (int n = 0; n < NP; ++n) y[n] = o.mul(splat(7.0376836292e-2f), m[n]);
#define UPS_PK_STEP(c)                                              \
    _Pragma("unroll") for (int n = 0; n < NP; ++n) y[n] = o.add(y[n], splat(c)); \
    _Pragma("unroll") for (int n = 0; n < NP; ++n) y[n] = o.mul(y[n], m[n]);
    UPS_PK_STEP(-1.1514610310e-1f)
    UPS_PK_STEP(1.1676998740e-1f)
    UPS_PK_STEP(-1.2420140846e-1f)
    UPS_PK_STEP(1.4249322787e-1f)
    UPS_PK_STEP(-1.6668057665e-1f)
    UPS_PK_STEP(2.0000714765e-1f)
    UPS_PK_STEP(-2.4999993993e-1f)
    UPS_PK_STEP(3.3333331174e-1f)   // ends with y = fl(fl(y*m + c8) * m)
#undef UPS_PK_STEP
#pragma unroll
    for (int n = 0; n < NP; ++n) y[n] = o.mul(y[n], z[n]);
#pragma unroll
    for (int n = 0; n < NP; ++n) y[n] = o.add(y[n], o.mul(ef[n], splat(-2.12194440e-4f)));
#pragma unroll
    for (int n = 0; n < NP; ++n) y[n] = o.sub(y[n], o.mul(splat(0.5f), z[n]));
#pragma unroll
    for (int n = 0; n < NP; ++n) out[n] = o.add(o.add(m[n], y[n]), o.mul(ef[n], splat(0.693359375f)));
}

__device__ __forceinline__ f2 log_canon2(const Ops& o, float x0, float x1) {
    const f2 x[1] = {pack(x0, x1)};
    f2 r[1];
    log_canon2n<1>(o, x, r);
    return r[0];
}

// exp_canon of 2*NP arguments, NP independent pairs in lockstep (see log_canon2n);
// lane results == canon_math.cuh::exp_canon
template <int NP>
__device__ __forceinline__ void exp_canon2n(const Ops& o, const f2 (&xin)[NP], f2 (&out)[NP]) {
    f2 x[NP], n[NP], r[NP], z[NP], y[NP], t[NP];
    bool zero0[NP], zero1[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        float x0, x1;
        unpack(xin[i], x0, x1);
        zero0[i] = x0 < -80.0f; zero1[i] = x1 < -80.0f;
        x0 = x0 > 88.0f ? 88.0f : x0; x1 = x1 > 88.0f ? 88.0f : x1;
        x0 = x0 < -80.0f ? -80.0f : x0; x1 = x1 < -80.0f ? -80.0f : x1;
        x[i] = pack(x0, x1);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) t[i] = o.mul(x[i], splat(1.44269504088896341f));
#pragma unroll
    for (int i = 0; i < NP; ++i) t[i] = o.add(t[i], splat(0.5f));
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        float t0, t1;
        unpack(t[i], t0, t1);
        n[i] = pack(floorf(t0), floorf(t1));
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) t[i] = o.mul(n[i], splat(0.693359375f));
#pragma unroll
    for (int i = 0; i < NP; ++i) r[i] = o.sub(x[i], t[i]);
#pragma unroll
    for (int i = 0; i < NP; ++i) t[i] = o.mul(n[i], splat(-2.12194440e-4f));
#pragma unroll
    for (int i = 0; i < NP; ++i) r[i] = o.sub(r[i], t[i]);
#pragma unroll
    for (int i = 0; i < NP; ++i) z[i] = o.mul(r[i], r[i]);
#pragma unroll
    for (int i = 0; i < NP; ++i) y[i] = o.mul(splat(1.9875691500e-4f), r[i]);
#define UPS_PK_STEP(c)                                                          \
    _Pragma("unroll") for (int i = 0; i < NP; ++i) y[i] = o.add(y[i], splat(c)); \
    _Pragma("unroll") for (int i = 0; i < NP; ++i) y[i] = o.mul(y[i], r[i]);
    UPS_PK_STEP(1.3981999507e-3f)
    UPS_PK_STEP(8.3334519073e-3f)
    UPS_PK_STEP(4.1665795894e-2f)
    UPS_PK_STEP(1.6666665459e-1f)
#undef UPS_PK_STEP
#pragma unroll
    for (int i = 0; i < NP; ++i) y[i] = o.add(y[i], splat(5.0000001201e-1f));
#pragma unroll
    for (int i = 0; i < NP; ++i) y[i] = o.mul(y[i], z[i]);
#pragma unroll
    for (int i = 0; i < NP; ++i) y[i] = o.add(y[i], r[i]);
#pragma unroll
    for (int i = 0; i < NP; ++i) y[i] = o.add(y[i], o.one);
    // 2^n: n is an integer in [-116, 127]; bits(n + 1.5*2^23) << 23 == n << 23 (mod 2^32)
#pragma unroll
    for (int i = 0; i < NP; ++i) t[i] = o.add(n[i], splat(MAGIC));
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        float u0, u1;
        unpack(t[i], u0, u1);
        const float s0 = __int_as_float((__float_as_int(u0) << 23) + 0x3F800000);
        const float s1 = __int_as_float((__float_as_int(u1) << 23) + 0x3F800000);
        float r0, r1;
        unpack(o.mul(y[i], pack(s0, s1)), r0, r1);
        out[i] = pack(zero0[i] ? 0.0f : r0, zero1[i] ? 0.0f : r1);
    }
}

__device__ __forceinline__ f2 exp_canon2(const Ops& o, float x0, float x1) {
    const f2 x[1] = {pack(x0, x1)};
    f2 r[1];
    exp_canon2n<1>(o, x, r);
    return r[0];
}

}  // namespace pk
}  // namespace ups
#endif
