// philox.cuh -- stateless Philox4x32-10 stream + Box-Muller normals, sm_100a device code.
//
// Replaces the reference's stateful XORWOW generator: setup_kernel / curand_init(seed, tid, 0)
// (inc/tool.cuh:192-195) and curand_normal(&state) at every call site
// (inc/trajectories.cuh:74,145,223,301; inc/nmc.cuh:56,152,223,336; inc/testing.cuh:67).
// The integer words are bit-identical to cuRAND's curandStatePhilox4_32_10_t stream
// curand_init(seed, subsequence, offset = 0): block b of subsequence p is
// Philox(ctr = (b_lo, b_hi, p_lo, p_hi), key = (seed_lo, seed_hi)).  No state is stored:
// the counter is a function of (path id, step), so there is no setup kernel and no 48-byte
// per-thread state traffic.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mcb {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

// The ten round keys are thread-invariant: they are expanded once on the host and live in
// the kernel's constant bank, so the key schedule costs no instructions on the device.
struct PhiloxKeys {
    uint32_t k0[10];
    uint32_t k1[10];
};

inline PhiloxKeys make_philox_keys(uint64_t seed)
{
    PhiloxKeys k;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int i = 0; i < 10; ++i) {
        k.k0[i] = a;
        k.k1[i] = b;
        a += kPhiloxW0;
        b += kPhiloxW1;
    }
    return k;
}

struct Words4 {
    uint32_t x, y, z, w;
};

// One round: 2 IMAD.WIDE.U32 + 2 LOP3 (three-input xor).
__device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3,
                                             uint32_t k0, uint32_t k1)
{
    const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;
    const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
}

// Full bijection.  When the caller passes literal zeros for the block counter (single-step
// pricing: block 0) the first round's M0 product folds away at compile time, the second
// round's M1 product is loop-invariant (it depends only on p_hi and the key), and unused
// output words are dead-code-eliminated round by round.
__device__ __forceinline__ Words4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                const PhiloxKeys &k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) philox_round(c0, c1, c2, c3, k.k0[r], k.k1[r]);
    return Words4{c0, c1, c2, c3};
}

// ---- Box-Muller on the MUFU (XU) pipe -------------------------------------------------
// cuRAND's _curand_box_muller (curand_normal.h:70-88): u = x*2^-32 + 2^-33 in (0,1],
// v = y*2pi*2^-32 + half a step; normal pair (s sin v, s cos v), s = sqrt(-2 ln u).
// Here ln comes from MUFU.LG2, the root from MUFU.SQRT and sin/cos from MUFU.SIN/COS.
// The angle word is converted as a SIGNED integer: (int)y*2pi*2^-32 differs from cuRAND's
// unsigned form by exactly 2pi when y >= 2^31, i.e. it is the same angle, but it lands in
// [-pi, pi) where MUFU.SIN/COS are most accurate (|err| <= 2^-21.4).

__device__ __forceinline__ float mufu_lg2(float x)
{
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float mufu_ex2(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float mufu_sqrt(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float mufu_sin(float x)
{
    float r;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float mufu_cos(float x)
{
    float r;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

constexpr float k2Pow32Inv = 2.3283064e-10f;                 // CURAND_2POW32_INV
constexpr float k2Pow32Inv2Pi = 2.3283064e-10f * 6.2831855f; // CURAND_2POW32_INV_2PI
constexpr float kSqrt2Ln2 = 1.1774100225154747f;             // sqrt(2 ln 2)

// Unscaled Box-Muller radius t = sqrt(-log2 u), so that s = sqrt(-2 ln u) = sqrt(2 ln 2) * t.
// The constant is never applied on its own: callers fold it into the scale they multiply the
// normal by anyway (sigma sqrt(dt) log2 e ...), which removes one FMUL per pair; the negation
// rides on the MUFU.SQRT operand.  I2FP, FFMA, MUFU.LG2, MUFU.SQRT.
__device__ __forceinline__ float bm_radius_unscaled(uint32_t x)
{
    const float u = fmaf(__uint2float_rn(x), k2Pow32Inv, 0.5f * k2Pow32Inv);
    return mufu_sqrt(-mufu_lg2(u));
}

// v in [-pi, pi) : I2FP, FFMA
__device__ __forceinline__ float bm_angle(uint32_t y)
{
    return fmaf(__int2float_rn((int)y), k2Pow32Inv2Pi, 0.5f * k2Pow32Inv2Pi);
}

// normal 0 of a Philox block (the even, "sin" member of the first pair) divided by sqrt(2 ln 2)
__device__ __forceinline__ float unit_normal_sin(uint32_t x, uint32_t y)
{
    return bm_radius_unscaled(x) * mufu_sin(bm_angle(y));
}

// The four normals of a Philox block, in curand_normal order, as affine images
//   d[j] = scale * normal_j / sqrt(2 ln 2) + shift.
// Multi-step walks pass scale = sigma sqrt(dt) log2(e) sqrt(2 ln 2) and shift = the per-step
// drift in log2 units, so d[j] IS the log2-price increment of step j and a step costs one FADD;
// scale = sqrt(2 ln 2), shift = 0 gives the plain normals.
__device__ __forceinline__ void increments4(const Words4 &w, float scale, float shift, float d[4])
{
    const float t0 = bm_radius_unscaled(w.x) * scale, v0 = bm_angle(w.y);
    const float t1 = bm_radius_unscaled(w.z) * scale, v1 = bm_angle(w.w);
    d[0] = fmaf(t0, mufu_sin(v0), shift);
    d[1] = fmaf(t0, mufu_cos(v0), shift);
    d[2] = fmaf(t1, mufu_sin(v1), shift);
    d[3] = fmaf(t1, mufu_cos(v1), shift);
}

// The four UNIT normals of a Philox block as (radius, trig) factors, left unmultiplied: normal j is
// r[j / 2] * g[j] with r the unscaled Box-Muller radius sqrt(-log2 u) and g = sin, cos, sin, cos.  The
// multi-step walks feed them straight into the accumulating FFMA (acc = fma(r, g, acc)): no scaling
// multiply and no separate add per step.
__device__ __forceinline__ void unit_factors4(const Words4 &w, float r[2], float g[4])
{
    r[0] = bm_radius_unscaled(w.x);
    r[1] = bm_radius_unscaled(w.z);
    const float v0 = bm_angle(w.y), v1 = bm_angle(w.w);
    g[0] = mufu_sin(v0);
    g[1] = mufu_cos(v0);
    g[2] = mufu_sin(v1);
    g[3] = mufu_cos(v1);
}

// ---- packed FP32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2) -------------------------------------
// One instruction works on TWO floats held in an even-aligned register pair.  The FMA pipes do
// not get faster (128 lanes per SM either way), but the hot kernels here are bound by ISSUE SLOTS
// (one warp instruction per clock per scheduler), and a packed instruction takes one slot for two
// operations.  Each half rounds exactly like the scalar instruction (round-to-nearest, no flush),
// so the packed and the scalar forms below are bit-identical.
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_sub(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// increments4 with the two Box-Muller pairs of a block worked side by side: 5 packed instructions
// (2 FFMA2 uniform maps, 1 FMUL2 radius scale, 2 FFMA2 increments) instead of 10 scalar ones.
// Same operations on the same operands as increments4 => the same bits.
__device__ __forceinline__ void increments4_packed(const Words4 &w, float scale, float shift, float d[4])
{
    const uint64_t uu = f2_fma(f2_pack(__uint2float_rn(w.x), __uint2float_rn(w.z)), f2_pack(k2Pow32Inv, k2Pow32Inv),
                               f2_pack(0.5f * k2Pow32Inv, 0.5f * k2Pow32Inv));
    const uint64_t vv = f2_fma(f2_pack(__int2float_rn((int)w.y), __int2float_rn((int)w.w)),
                               f2_pack(k2Pow32Inv2Pi, k2Pow32Inv2Pi),
                               f2_pack(0.5f * k2Pow32Inv2Pi, 0.5f * k2Pow32Inv2Pi));
    float u0, u1, v0, v1;
    f2_unpack(uu, u0, u1);
    f2_unpack(vv, v0, v1);
    const uint64_t tt = f2_mul(f2_pack(mufu_sqrt(-mufu_lg2(u0)), mufu_sqrt(-mufu_lg2(u1))), f2_pack(scale, scale));
    float t0, t1;
    f2_unpack(tt, t0, t1);
    const uint64_t sh = f2_pack(shift, shift);
    f2_unpack(f2_fma(f2_pack(t0, t0), f2_pack(mufu_sin(v0), mufu_cos(v0)), sh), d[0], d[1]);
    f2_unpack(f2_fma(f2_pack(t1, t1), f2_pack(mufu_sin(v1), mufu_cos(v1)), sh), d[2], d[3]);
}

}  // namespace mcb
