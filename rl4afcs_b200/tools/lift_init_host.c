/* BUILD TOOL (host side of tools/lift_plant.py) -- runs the translated initialize() of the plant binary once, at translation
 * time, over a flat copy of the DLL image and writes the resulting image to stdout.  The translator uses it for its second pass:
 * everything outside the regions step() writes is CONSTANT while aircraft are stepping, so loads from it can be folded into
 * the generated code (pointer chains through the Simulink SimStruct, block parameters, table addresses).
 *
 *     gcc -O1 -ffp-contract=off -DLIFT_GENERATED_INC='"..._code_init.inc"' lift_init_host.c -lm && ./a.out image.bin > post_init.bin
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MEM_SIZE 0x60000ULL            /* image 0x40000 + stack */
#define LIFT_CPU_EXTRA uint8_t* M;
#include "../../include/rl4_lift_runtime.h"

static void trap(const char* msg, uint64_t v) { fprintf(stderr, "lift_init_host: %s (0x%llx)\n", msg, (unsigned long long)v); exit(3); }
#define LIFT_TRAP(msg, v) do { trap(msg, (uint64_t)(v)); LIFT_TRAP_RETURN; } while (0)
static inline uint8_t* at(cpu_t* c, uint32_t a32, unsigned n)
{
    const uint32_t off = a32 - (uint32_t)LIFT_BASE;
    if (off > MEM_SIZE - n) trap("access outside the emulated address space", a32);
    return c->M + off;
}
#define LD_(T, a) ({ T v_; memcpy(&v_, at(c, (uint32_t)(a), sizeof(T)), sizeof(T)); v_; })
#define ST_(T, a, v) do { T v_ = (T)(v); memcpy(at(c, (uint32_t)(a), sizeof(T)), &v_, sizeof(T)); } while (0)
#define LD8(a) ((uint64_t)LD_(uint8_t, a))
#define LD16(a) ((uint64_t)LD_(uint16_t, a))
#define LD32(a) ((uint64_t)LD_(uint32_t, a))
#define LD64(a) LD_(uint64_t, a)
#define LDD(a) LD_(double, a)
#define ST8(a, v) ST_(uint8_t, a, v)
#define ST16(a, v) ST_(uint16_t, a, v)
#define ST32(a, v) ST_(uint32_t, a, v)
#define ST64(a, v) ST_(uint64_t, a, v)
#define LDS8 LD8
#define LDS16 LD16
#define LDS32 LD32
#define LDS64 LD64
#define LDSD LDD
#define STS8 ST8
#define STS16 ST16
#define STS32 ST32
#define STS64 ST64
#define LDW8 LD8
#define LDW16 LD16
#define LDW32 LD32
#define LDW64 LD64
#define LDWD LDD
#define STW8 ST8
#define STW16 ST16
#define STW32 ST32
#define STW64 ST64
#define LDI8 LD8
#define LDI16 LD16
#define LDI32 LD32
#define LDI64 LD64
#define LDID LDD
#define lift_cos cos
#define lift_sin sin
#define lift_tan tan
#define lift_exp exp
#define lift_floor floor
#define lift_log10 log10
#define lift_sqrt sqrt
#define lift_pow pow
static void lift_memcpy(cpu_t* c, uint64_t d, uint64_t s, uint64_t n) { if (n) memmove(at(c, (uint32_t)d, (unsigned)n), at(c, (uint32_t)s, (unsigned)n), n); }
static void lift_memset(cpu_t* c, uint64_t d, int v, uint64_t n) { if (n) memset(at(c, (uint32_t)d, (unsigned)n), v, n); }
static uint64_t lift_malloc(cpu_t* c, uint64_t n) { (void)c; trap("the model allocates memory", n); return 0; }
static void lift_REPSTOS(cpu_t* c, uint64_t* rcx, uint64_t* rdi, uint64_t rax, unsigned w)
{
    for (; *rcx; --*rcx, *rdi += w) memcpy(at(c, (uint32_t)*rdi, w), &rax, w);
}
static void lift_REPMOVS(cpu_t* c, uint64_t* rcx, uint64_t* rdi, uint64_t* rsi, unsigned w)
{
    for (; *rcx; --*rcx, *rdi += w, *rsi += w) memmove(at(c, (uint32_t)*rdi, w), at(c, (uint32_t)*rsi, w), w);
}
static inline uint64_t lift_CVTR32(double v) { return (v > -2147483649.0 && v < 2147483648.0) ? (uint64_t)(uint32_t)(int32_t)nearbyint(v) : 0x80000000ULL; }
static inline uint64_t lift_CVTR64(double v) { return (uint64_t)(int64_t)nearbyint(v); }

#include LIFT_GENERATED_INC

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    static uint8_t mem[MEM_SIZE];
    const size_t n = fread(mem, 1, 0x40000, f);
    fclose(f);
    if (n == 0) return 2;
    cpu_t c;
    memset(&c, 0, sizeof c);
    c.M = mem;
    c.r[4] = LIFT_BASE + MEM_SIZE - 0x100 - 8;
    LIFT_INVOKE(f_1800096f0, &c);
    fwrite(mem, 1, 0x40000, stdout);
    return 0;
}
