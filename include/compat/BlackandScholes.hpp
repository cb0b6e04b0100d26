// compat/BlackandScholes.hpp -- drop-in for inc/BlackandScholes.hpp:8-43 (CND, black_scholes_CPU).
// Same signatures and the same arithmetic recipe (Abramowitz-Stegun 26.2.17 five-term polynomial
// in float; d1 with a double 0.5; float exp for the discount) so the printed closed form matches
// the reference's digit for digit.  tests/test_compat_headers.py pins it against the fixture
// generated from the unmodified reference (tests/golden/reference_cpu.json).
#ifndef MCB_COMPAT_BLACK_SCHOLES_HPP
#define MCB_COMPAT_BLACK_SCHOLES_HPP

#include <cmath>

inline float CND(float x)
{
    static const float coeff[5] = {0.31938153f, -0.356563782f, 1.781477937f, -1.821255978f, 1.330274429f};
    const float magnitude = x < 0.0f ? -x : x;
    const float t = 1.0f / (1.0f + 0.2316419f * magnitude);
    float horner = coeff[4];
    for (int i = 3; i >= 0; --i) horner = horner * t + coeff[i];
    const float upper_tail = 0.39894228f * expf(-x * x / 2.0f) * t * horner;
    return x < 0.0f ? upper_tail : 1.0f - upper_tail;
}

inline void black_scholes_CPU(float &callResult, float S0, float K, float T, float r, float v)
{
    const float root_t = sqrtf(T);
    const float d1 = (float)((logf(S0 / K) + (r + 0.5 * v * v) * T) / (v * root_t));
    const float d2 = d1 - v * root_t;
    callResult = S0 * CND(d1) - K * expf(-r * T) * CND(d2);
}

#endif
