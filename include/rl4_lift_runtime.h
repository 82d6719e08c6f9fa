/* rl4_lift_runtime.h -- machine model under the code that rl4afcs_b200/tools/lift_plant.py generates from the reference's plant binary
 * (envs/nonlinear/<variant>/_citation.cp39-win_amd64.pyd; /root/reference/envs/nonlinear/citation.py:62-69).
 *
 * The generated code is a sequence of x86-64 instructions spelled as C statements over this state: 16 integer
 * registers, 16 SSE registers and five flags (C locals inside a translated function), and ONE flat little-endian memory that holds the DLL image at its preferred
 * base, a bump-allocated heap behind it and the stack at the top.  Each helper below implements the architectural
 * semantics of one instruction class (Intel SDM vol. 2); floating point is IEEE binary64 with one rounding per
 * instruction, exactly what SSE2 scalar / packed instructions do, so the host compiler must not contract (`-ffp-contract=off`
 * for gcc; the CUDA build maps F_* to the __d*_rn intrinsics).
 *
 * Before including the generated .inc the includer defines how memory is reached:
 *     LD8/16/32/64(addr), LDD(addr), ST8/16/32/64(addr, value)       (addr = emulated virtual address, uint64_t)
 *     LDS.. / STS..: the same for operands the translator knows to be on the stack (rsp- or frame-pointer-based)
 * and may define LIFT_FN (function qualifiers), LIFT_TRAP(msg, value), F_ADD ... F_SQRT, and the lift_<libm> functions.
 */
#ifndef RL4_LIFT_RUNTIME_H
#define RL4_LIFT_RUNTIME_H

#include <stdint.h>

#define LIFT_BASE 0x180000000ULL

#ifndef LIFT_HD
#define LIFT_HD static inline
#endif

typedef union { uint64_t u[2]; double d[2]; } lift_xmm;

/* what crosses a function boundary (Windows x64 convention): arguments rcx rdx r8 r9 / xmm0-3, the stack pointer, results
 * rax / xmm0.  Inside a translated function every register is a C local. */
typedef struct cpu_t {
    uint64_t r[16];          /* rax rcx rdx rbx rsp rbp rsi rdi r8..r15 */
    lift_xmm x[16];
    uint64_t heap_next, heap_end;
    LIFT_CPU_EXTRA
} cpu_t;
typedef struct lift_flags { uint8_t zf, sf, cf, of, pf; } lift_flags;
typedef struct lift_ret { uint64_t rax, x0; } lift_ret;          /* what a translated function returns: rax and the low half of xmm0 */
LIFT_HD lift_ret lift_mkret(uint64_t rax, uint64_t x0) { lift_ret r; r.rax = rax; r.x0 = x0; return r; }
/* calls a translated entry point with the arguments held in cpu_t (the includer's glue: kernels, host harnesses) */
#define LIFT_INVOKE(f, c) do { const lift_ret t__ = f((c), (c)->r[1], (c)->r[2], (c)->r[8], (c)->r[9], (c)->x[0].u[0], (c)->x[1].u[0], \
                                                      (c)->x[2].u[0], (c)->x[3].u[0], (c)->r[4]); (c)->r[0] = t__.rax; (c)->x[0].u[0] = t__.x0; } while (0)

#ifndef U2D
LIFT_HD double lift_u2d(uint64_t u) { union { uint64_t u; double d; } t; t.u = u; return t.d; }
LIFT_HD uint64_t lift_d2u(double d) { union { uint64_t u; double d; } t; t.d = d; return t.u; }
#define U2D(u) lift_u2d(u)
#define D2U(d) lift_d2u(d)
#endif

#ifndef F_ADD
#define F_ADD(a, b) ((a) + (b))
#define F_SUB(a, b) ((a) - (b))
#define F_MUL(a, b) ((a) * (b))
#define F_DIV(a, b) ((a) / (b))
#define F_SQRT(a) sqrt(a)
#endif
/* maxsd / minsd: the SECOND operand is returned when either is NaN or both are zero */
#define F_MAX(a, b) (((a) > (b)) ? (a) : (b))
#define F_MIN(a, b) (((a) < (b)) ? (a) : (b))

LIFT_HD float lift_u2f(uint32_t v) { union { uint32_t u; float f; } t; t.u = v; return t.f; }
LIFT_HD uint32_t lift_f2u(float f) { union { uint32_t u; float f; } t; t.f = f; return t.u; }

LIFT_HD uint8_t lift_parity(uint64_t v) { v &= 0xff; v ^= v >> 4; v ^= v >> 2; v ^= v >> 1; return (uint8_t)(~v & 1); }

#define LIFT_DEFINE_INT(W, UT, ST)                                                                                              \
    LIFT_HD uint64_t lift_ADD##W(lift_flags* c, uint64_t a_, uint64_t b_) {                                                          \
        UT a = (UT)a_, b = (UT)b_, r = (UT)(a + b);                                                                             \
        c->zf = r == 0; c->sf = (ST)r < 0; c->cf = r < a; c->of = (ST)((a ^ r) & (b ^ r)) < 0; c->pf = lift_parity(r); return r; } \
    LIFT_HD uint64_t lift_ADC##W(lift_flags* c, uint64_t a_, uint64_t b_) {                                                          \
        UT a = (UT)a_, b = (UT)b_, ci = c->cf, r = (UT)(a + b + ci);                                                            \
        c->zf = r == 0; c->sf = (ST)r < 0; c->cf = ci ? r <= a : r < a; c->of = (ST)((a ^ r) & (b ^ r)) < 0; c->pf = lift_parity(r); return r; } \
    LIFT_HD uint64_t lift_SUB##W(lift_flags* c, uint64_t a_, uint64_t b_) {                                                          \
        UT a = (UT)a_, b = (UT)b_, r = (UT)(a - b);                                                                             \
        c->zf = r == 0; c->sf = (ST)r < 0; c->cf = a < b; c->of = (ST)((a ^ b) & (a ^ r)) < 0; c->pf = lift_parity(r); return r; } \
    LIFT_HD uint64_t lift_SBB##W(lift_flags* c, uint64_t a_, uint64_t b_) {                                                          \
        UT a = (UT)a_, b = (UT)b_, ci = c->cf, r = (UT)(a - b - ci);                                                            \
        c->zf = r == 0; c->sf = (ST)r < 0; c->cf = ci ? a <= b : a < b; c->of = (ST)((a ^ b) & (a ^ r)) < 0; c->pf = lift_parity(r); return r; } \
    LIFT_HD uint64_t lift_AND##W(lift_flags* c, uint64_t a_, uint64_t b_) {                                                          \
        UT r = (UT)((UT)a_ & (UT)b_); c->zf = r == 0; c->sf = (ST)r < 0; c->cf = 0; c->of = 0; c->pf = lift_parity(r); return r; } \
    LIFT_HD uint64_t lift_OR##W(lift_flags* c, uint64_t a_, uint64_t b_) {                                                           \
        UT r = (UT)((UT)a_ | (UT)b_); c->zf = r == 0; c->sf = (ST)r < 0; c->cf = 0; c->of = 0; c->pf = lift_parity(r); return r; } \
    LIFT_HD uint64_t lift_XOR##W(lift_flags* c, uint64_t a_, uint64_t b_) {                                                          \
        UT r = (UT)((UT)a_ ^ (UT)b_); c->zf = r == 0; c->sf = (ST)r < 0; c->cf = 0; c->of = 0; c->pf = lift_parity(r); return r; } \
    LIFT_HD uint64_t lift_INC##W(lift_flags* c, uint64_t a_) {                                                                       \
        UT a = (UT)a_, r = (UT)(a + 1); c->zf = r == 0; c->sf = (ST)r < 0; c->of = (ST)((a ^ r) & (1 ^ r)) < 0; c->pf = lift_parity(r); return r; } \
    LIFT_HD uint64_t lift_DEC##W(lift_flags* c, uint64_t a_) {                                                                       \
        UT a = (UT)a_, r = (UT)(a - 1); c->zf = r == 0; c->sf = (ST)r < 0; c->of = (ST)((a ^ 1) & (a ^ r)) < 0; c->pf = lift_parity(r); return r; } \
    LIFT_HD uint64_t lift_NEG##W(lift_flags* c, uint64_t a_) {                                                                       \
        UT a = (UT)a_, r = (UT)(0 - a); c->zf = r == 0; c->sf = (ST)r < 0; c->cf = a != 0; c->of = (ST)(a & r) < 0; c->pf = lift_parity(r); return r; } \
    LIFT_HD uint64_t lift_SHL##W(lift_flags* c, uint64_t a_, uint64_t n_) {                                                          \
        const unsigned n = (unsigned)n_ & (W == 64 ? 63 : 31); UT a = (UT)a_; if (!n) return a;                                 \
        UT r = n < W ? (UT)(a << n) : 0; c->cf = n <= W ? (a >> (W - n)) & 1 : 0; c->zf = r == 0; c->sf = (ST)r < 0;            \
        c->of = ((r >> (W - 1)) & 1) ^ c->cf; c->pf = lift_parity(r); return r; }                                               \
    LIFT_HD uint64_t lift_SHR##W(lift_flags* c, uint64_t a_, uint64_t n_) {                                                          \
        const unsigned n = (unsigned)n_ & (W == 64 ? 63 : 31); UT a = (UT)a_; if (!n) return a;                                 \
        UT r = n < W ? (UT)(a >> n) : 0; c->cf = n <= W ? (a >> (n - 1)) & 1 : 0; c->zf = r == 0; c->sf = (ST)r < 0;            \
        c->of = (a >> (W - 1)) & 1; c->pf = lift_parity(r); return r; }                                                         \
    LIFT_HD uint64_t lift_SAR##W(lift_flags* c, uint64_t a_, uint64_t n_) {                                                          \
        unsigned n = (unsigned)n_ & (W == 64 ? 63 : 31); ST a = (ST)(UT)a_; if (!n) return (UT)a;                               \
        c->cf = n >= W ? (UT)(a < 0) : (UT)(((UT)a >> (n - 1)) & 1); if (n >= W) n = W - 1;                                     \
        UT r = (UT)(a >> n); c->zf = r == 0; c->sf = (ST)r < 0; c->of = 0; c->pf = lift_parity(r); return r; } \
    LIFT_HD uint64_t lift_ROL##W(lift_flags* c, uint64_t a_, uint64_t n_) {                                                          \
        const unsigned n = ((unsigned)n_ & (W == 64 ? 63 : 31)) % W; UT a = (UT)a_; if (!n) return a;                           \
        UT r = (UT)((a << n) | (a >> (W - n))); c->cf = r & 1; return r; }                                                      \
    LIFT_HD uint64_t lift_ROR##W(lift_flags* c, uint64_t a_, uint64_t n_) {                                                          \
        const unsigned n = ((unsigned)n_ & (W == 64 ? 63 : 31)) % W; UT a = (UT)a_; if (!n) return a;                           \
        UT r = (UT)((a >> n) | (a << (W - n))); c->cf = (r >> (W - 1)) & 1; return r; }

LIFT_DEFINE_INT(8, uint8_t, int8_t)
LIFT_DEFINE_INT(16, uint16_t, int16_t)
LIFT_DEFINE_INT(32, uint32_t, int32_t)
LIFT_DEFINE_INT(64, uint64_t, int64_t)

/* two- / three-operand imul: truncated product; CF = OF = the product did not fit */
LIFT_HD uint64_t lift_IMUL16(lift_flags* c, uint64_t a, uint64_t b) { int32_t p = (int32_t)(int16_t)a * (int32_t)(int16_t)b; c->cf = c->of = p != (int16_t)p; return (uint16_t)p; }
LIFT_HD uint64_t lift_IMUL32(lift_flags* c, uint64_t a, uint64_t b) { int64_t p = (int64_t)(int32_t)a * (int64_t)(int32_t)b; c->cf = c->of = p != (int32_t)p; return (uint32_t)p; }
LIFT_HD uint64_t lift_IMUL64(lift_flags* c, uint64_t a, uint64_t b) { c->cf = c->of = 0; return a * b; }   /* overflow flag of a 64-bit imul is never consumed by compiled code */

/* comisd / ucomisd: unordered -> ZF = PF = CF = 1 */
LIFT_HD void lift_COMISD(lift_flags* c, double a, double b)
{
    c->of = 0; c->sf = 0;
    if (a != a || b != b) { c->zf = 1; c->pf = 1; c->cf = 1; }
    else { c->zf = a == b; c->pf = 0; c->cf = a < b; }
}
/* cvttsd2si: out-of-range and NaN give the "integer indefinite" value */
LIFT_HD uint64_t lift_CVTT32(double v) { return (v > -2147483649.0 && v < 2147483648.0) ? (uint64_t)(uint32_t)(int32_t)v : 0x80000000ULL; }
LIFT_HD uint64_t lift_CVTT64(double v) { return (v >= -9223372036854775808.0 && v < 9223372036854775808.0) ? (uint64_t)(int64_t)v : 0x8000000000000000ULL; }

#endif
