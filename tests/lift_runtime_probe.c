/* test helper: exposes the x86 instruction-semantics helpers of include/rl4_lift_runtime.h (the machine model under the
 * translated plant binary) so that tests/test_lift_runtime.py can check them against an independent model of the ISA */
#include <math.h>
#include <stdint.h>
#define LIFT_CPU_EXTRA int unused;
#include "../include/rl4_lift_runtime.h"

#define PROBE2(NAME, W)                                                                               \
    uint64_t probe_##NAME##W(uint64_t a, uint64_t b, int cf_in, uint8_t* flags)                        \
    {                                                                                                  \
        lift_flags f = {0, 0, 0, 0, 0};                                                                \
        f.cf = (uint8_t)cf_in;                                                                         \
        const uint64_t r = lift_##NAME##W(&f, a, b);                                                   \
        flags[0] = f.zf; flags[1] = f.sf; flags[2] = f.cf; flags[3] = f.of; flags[4] = f.pf;           \
        return r;                                                                                      \
    }
#define PROBE1(NAME, W)                                                                               \
    uint64_t probe_##NAME##W(uint64_t a, uint64_t b, int cf_in, uint8_t* flags)                        \
    {                                                                                                  \
        lift_flags f = {0, 0, 0, 0, 0};                                                                \
        (void)b; f.cf = (uint8_t)cf_in;                                                                \
        const uint64_t r = lift_##NAME##W(&f, a);                                                      \
        flags[0] = f.zf; flags[1] = f.sf; flags[2] = f.cf; flags[3] = f.of; flags[4] = f.pf;           \
        return r;                                                                                      \
    }
#define ALL(W) PROBE2(ADD, W) PROBE2(ADC, W) PROBE2(SUB, W) PROBE2(SBB, W) PROBE2(AND, W) PROBE2(OR, W) PROBE2(XOR, W) \
               PROBE1(INC, W) PROBE1(DEC, W) PROBE1(NEG, W) PROBE2(SHL, W) PROBE2(SHR, W) PROBE2(SAR, W) PROBE2(ROL, W) PROBE2(ROR, W)
ALL(8) ALL(16) ALL(32) ALL(64)
PROBE2(IMUL, 32) PROBE2(IMUL, 64)

void probe_comisd(double a, double b, uint8_t* flags)
{
    lift_flags f = {9, 9, 9, 9, 9};
    lift_COMISD(&f, a, b);
    flags[0] = f.zf; flags[1] = f.sf; flags[2] = f.cf; flags[3] = f.of; flags[4] = f.pf;
}
uint64_t probe_cvtt32(double v) { return lift_CVTT32(v); }
uint64_t probe_cvtt64(double v) { return lift_CVTT64(v); }
double probe_max(double a, double b) { return F_MAX(a, b); }
double probe_min(double a, double b) { return F_MIN(a, b); }
