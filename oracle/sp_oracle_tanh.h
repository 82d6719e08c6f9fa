/* TEST INFRASTRUCTURE ONLY -- the "t13" tanh, restated for the CPU oracle.
 *
 * The reference calls np.tanh / tf.tanh (objects.py:127), neither of which is
 * bit-reproducible across hosts (np.tanh differs from libm by up
 * to 3 ulp).  To make everything ELSE on the path bit-comparable, DESIGN.md defines one
 * tanh algorithm in words; the CUDA kernels implement it (rl4afcs_b200/csrc/rl4_math.cuh)
 * and this file restates it independently for the oracle.  With IEEE-754 +,-,*,/,fma
 * and no contraction both produce identical bits.
 *
 *   ax = |x| ;  ax >= BIG  ->  copysign(1, x)   (NaN -> NaN)
 *   t  = ax + ax
 *   kd = fma(t, log2(e), MAGIC) ; n = kd - MAGIC        (round-to-nearest integer)
 *   r  = fma(-n, ln2_hi, t) ; r = fma(-n, ln2_lo, r)
 *   q  = Horner(1/D!, ..., 1/2!) in r        (D = 13 for f64, 8 for f32)
 *   p  = fma(r*r, q, r)                      (= expm1(r))
 *   s  = 2^n ; em = fma(s, p, s - 1)         (= expm1(2|x|))
 *   f32:  y = em / (em + 2)                  (IEEE division)
 *   f64:  y = em (/) (em + 2), a fixed sequence of IEEE operations:
 *           d  = em + 2 ; r0 = (double)(1.0f / (float)d)           (float32 reciprocal of the float32-rounded d)
 *           e  = fma(-d, r0, 1) ; e = fma(e, e, e) ; rc = fma(r0, e, r0)      (one cubic Newton step)
 *           q0 = em * rc ; y = fma(rc, fma(-d, q0, em), q0)                   (quotient, remainder, correction)
 *   return copysign(y, x)
 * The kernels evaluate the f64 polynomial on r/2 with power-of-two-scaled coefficients; scaling by powers of two
 * commutes with rounding, so that is the same function bit for bit -- this file keeps the textbook form.
 */
#ifndef RL4_SP_ORACLE_TANH_H
#define RL4_SP_ORACLE_TANH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline double orc_t13_f64(double x)
{
    const double ax = fabs(x);
    if (!(ax < 19.0625)) {
        if (ax != ax) return x + x;
        return copysign(1.0, x);
    }
    const double t = ax + ax;
    const double MAGIC = 6755399441055744.0; /* 1.5 * 2^52 */
    const double kd = fma(t, 1.4426950408889634074, MAGIC);
    const double n = kd - MAGIC;
    double r = fma(-n, 6.93147180559945286227e-01, t);
    r = fma(-n, 2.31904681384629955842e-17, r);
    double q = 1.0 / 6227020800.0;          /* 1/13! */
    q = fma(q, r, 1.0 / 479001600.0);
    q = fma(q, r, 1.0 / 39916800.0);
    q = fma(q, r, 1.0 / 3628800.0);
    q = fma(q, r, 1.0 / 362880.0);
    q = fma(q, r, 1.0 / 40320.0);
    q = fma(q, r, 1.0 / 5040.0);
    q = fma(q, r, 1.0 / 720.0);
    q = fma(q, r, 1.0 / 120.0);
    q = fma(q, r, 1.0 / 24.0);
    q = fma(q, r, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    const double p = fma(r * r, q, r);
    uint64_t kb; memcpy(&kb, &kd, 8);
    const uint64_t sb = (uint64_t)(1023 + (int64_t)(int32_t)(uint32_t)kb) << 52;
    double s; memcpy(&s, &sb, 8);
    const double em = fma(s, p, s - 1.0);
    const double d = em + 2.0;
    const volatile float df = (float)d;      /* volatile: no excess precision, no reciprocal tricks */
    const volatile float rf = 1.0f / df;
    const double r0 = (double)rf;
    double e = fma(-d, r0, 1.0);
    e = fma(e, e, e);
    const double rc = fma(r0, e, r0);
    const double q0 = em * rc;
    const double y = fma(rc, fma(-d, q0, em), q0);
    return copysign(y, x);
}

static inline float orc_t13_f32(float x)
{
    const float ax = fabsf(x);
    if (!(ax < 9.125f)) {
        if (ax != ax) return x + x;
        return copysignf(1.0f, x);
    }
    const float t = ax + ax;
    const float MAGIC = 12582912.0f; /* 1.5 * 2^23 */
    const float kd = fmaf(t, 1.44269504088896341f, MAGIC);
    const float n = kd - MAGIC;
    float r = fmaf(-n, 6.93147182464599609375e-01f, t);
    r = fmaf(-n, -1.90465429995776804525e-09f, r);
    float q = 1.0f / 40320.0f;               /* 1/8! */
    q = fmaf(q, r, 1.0f / 5040.0f);
    q = fmaf(q, r, 1.0f / 720.0f);
    q = fmaf(q, r, 1.0f / 120.0f);
    q = fmaf(q, r, 1.0f / 24.0f);
    q = fmaf(q, r, 1.0f / 6.0f);
    q = fmaf(q, r, 0.5f);
    const float p = fmaf(r * r, q, r);
    uint32_t kb; memcpy(&kb, &kd, 4);
    const int32_t ni = (int32_t)(kb & 0x3fffffu) ;   /* kd = 1.5*2^23 + n, 0 <= n < 2^22 */
    const uint32_t sb = (uint32_t)(127 + ni) << 23;
    float s; memcpy(&s, &sb, 4);
    const float em = fmaf(s, p, s - 1.0f);
    const float y = em / (em + 2.0f);
    return copysignf(y, x);
}
#endif
