/* TEST INFRASTRUCTURE -- CPU build of the translated plant binary (rl4afcs_b200/tools/lift_plant.py) -> oracle/_ref/libcitation_lifted.so.
 *
 * Same three entry points as the reference's `citation` module (envs/nonlinear/citation.py:62-69), but re-entrant: every
 * instance owns its flat memory (DLL image + heap + stack), so any number of aircraft can be stepped side by side and the
 * library runs where the .pyd cannot (no x86 / no Windows ABI needed -- only this file, include/rl4_lift_runtime.h and the generated
 * .inc).  Checked bit for bit against the binary executing natively (pe_citation.c) by tests/test_citation_lifted.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define LIFT_IMAGE_SIZE 0x40000ULL
#define LIFT_HEAP_SIZE  0x40000ULL
#define LIFT_STACK_SIZE 0x10000ULL
#define LIFT_MEM_SIZE   (LIFT_IMAGE_SIZE + LIFT_HEAP_SIZE + LIFT_STACK_SIZE)

#define LIFT_CPU_EXTRA uint8_t* M; uint64_t n_ins; uint8_t* wmask; int guard;
#include "../../include/rl4_lift_runtime.h"

static void lift_trap(const char* msg, uint64_t v)
{
    fprintf(stderr, "citation_lifted: %s (0x%llx)\n", msg, (unsigned long long)v);
    abort();
}
#define LIFT_TRAP(msg, v) do { lift_trap(msg, (uint64_t)(v)); LIFT_TRAP_RETURN; } while (0)

/* memory operands arrive as the LOW 32 BITS of the emulated address (lift.py: mem_a32); helpers that take full 64-bit
 * addresses (memcpy & co.) truncate the same way */
static inline uint8_t* lift_ptr(cpu_t* c, uint64_t a, unsigned n)
{
    const uint64_t off = (uint32_t)((uint32_t)a - (uint32_t)LIFT_BASE);
    if (off > LIFT_MEM_SIZE - n) lift_trap("memory access outside the emulated address space", a);
    /* [0x40000, 0x40800) is the virtual range of the private frames of leaf functions (LIFT_LEAF_LO32): they live in C arrays,
     * so nothing may reach that range through memory -- a pointer into such a frame that the analysis missed would */
    if (off - 0x40000u < 0x800u) lift_trap("access to the virtual range of a private leaf frame through memory", a);
    return c->M + off;
}
static inline uint8_t* lift_wptr(cpu_t* c, uint64_t a, unsigned n)
{
    uint8_t* p = lift_ptr(c, a, n);
    if (c->guard) {
        /* after initialize(): the translator folded loads from these parts of the writable window (lift_plant.py FROZEN_GAPS) and
         * from the image outside the regions step() writes -- a store there would invalidate the folded code */
        const uint32_t off = (uint32_t)a - (uint32_t)LIFT_BASE;
        if (off < 0x3a078u ? !(off - 0x2eb00u < 0x100u) : (off - 0x3a4b8u < 0x88u || off - 0x3a574u < 0x74u || off - 0x3a5f0u < 0x590u))
            lift_trap("store into memory the translation assumes constant after initialize()", a);
    }
    if (c->wmask) memset(c->wmask + (uint32_t)((uint32_t)a - (uint32_t)LIFT_BASE), 1, n);
    return p;
}
#define LIFT_LD(T, a) ({ T v_; memcpy(&v_, lift_ptr(c, (a), sizeof(T)), sizeof(T)); v_; })
#define LD8(a)  ((uint64_t)LIFT_LD(uint8_t, a))
#define LD16(a) ((uint64_t)LIFT_LD(uint16_t, a))
#define LD32(a) ((uint64_t)LIFT_LD(uint32_t, a))
#define LD64(a) LIFT_LD(uint64_t, a)
#define LDD(a)  LIFT_LD(double, a)
#define LIFT_ST(T, a, v) do { T v_ = (T)(v); memcpy(lift_wptr(c, (a), sizeof(T)), &v_, sizeof(T)); } while (0)
#define ST8(a, v)  LIFT_ST(uint8_t, a, v)
#define ST16(a, v) LIFT_ST(uint16_t, a, v)
#define ST32(a, v) LIFT_ST(uint32_t, a, v)
#define ST64(a, v) LIFT_ST(uint64_t, a, v)
/* region hints of the translator (W: writable window of .data, I: read-only image): same flat memory here, but every hinted
 * access is CHECKED against the region it claims -- the CUDA build trusts the hints */
static inline uint32_t lift_chk(cpu_t* c, uint32_t a32, int window)
{
    const uint32_t off = a32 - (uint32_t)LIFT_BASE;
    (void)c;
    if (window ? !(off - 0x3a000u < 0x2200u) : !(off - 0x13000u < 0x27000u && !(off - 0x2eb00u < 0x100u)))
        lift_trap(window ? "an access hinted as 'window' left the window" : "an access hinted as 'read-only image' left it", a32);
    return a32;
}
#define LDW8(a) LD8(lift_chk(c, (a), 1))
#define LDW16(a) LD16(lift_chk(c, (a), 1))
#define LDW32(a) LD32(lift_chk(c, (a), 1))
#define LDW64(a) LD64(lift_chk(c, (a), 1))
#define LDWD(a) LDD(lift_chk(c, (a), 1))
#define STW8(a, v) ST8(lift_chk(c, (a), 1), v)
#define STW16(a, v) ST16(lift_chk(c, (a), 1), v)
#define STW32(a, v) ST32(lift_chk(c, (a), 1), v)
#define STW64(a, v) ST64(lift_chk(c, (a), 1), v)
#define LDI8(a) LD8(lift_chk(c, (a), 0))
#define LDI16(a) LD16(lift_chk(c, (a), 0))
#define LDI32(a) LD32(lift_chk(c, (a), 0))
#define LDI64(a) LD64(lift_chk(c, (a), 0))
#define LDID(a) LDD(lift_chk(c, (a), 0))
static inline uint32_t lift_chk_stack(uint32_t a32)
{
    if (a32 - (uint32_t)LIFT_BASE < LIFT_IMAGE_SIZE + LIFT_HEAP_SIZE) lift_trap("an access hinted as 'stack' is not on the stack", a32);
    return a32;
}
#define LDS8(a) LD8(lift_chk_stack(a))
#define LDS16(a) LD16(lift_chk_stack(a))
#define LDS32(a) LD32(lift_chk_stack(a))
#define LDS64(a) LD64(lift_chk_stack(a))
#define LDSD(a) LDD(lift_chk_stack(a))
#define STS8(a, v) ST8(lift_chk_stack(a), v)
#define STS16(a, v) ST16(lift_chk_stack(a), v)
#define STS32(a, v) ST32(lift_chk_stack(a), v)
#define STS64(a, v) ST64(lift_chk_stack(a), v)

#define lift_cos cos
#define lift_sin sin
#define lift_tan tan
#define lift_exp exp
#define lift_floor floor
#define lift_log10 log10
#define lift_sqrt sqrt
#define lift_pow pow

static void lift_memcpy(cpu_t* c, uint64_t d, uint64_t s, uint64_t n) { if (n) memmove(lift_wptr(c, d, (unsigned)n), lift_ptr(c, s, (unsigned)n), n); }
static void lift_memset(cpu_t* c, uint64_t d, int v, uint64_t n) { if (n) memset(lift_wptr(c, d, (unsigned)n), v, n); }
static uint64_t lift_malloc(cpu_t* c, uint64_t n)
{
    const uint64_t p = (c->heap_next + 15) & ~15ULL;
    if (p + n > c->heap_end) lift_trap("emulated heap exhausted", n);
    c->heap_next = p + n;
    return p;
}
static void lift_REPSTOS(cpu_t* c, uint64_t* rcx, uint64_t* rdi, uint64_t rax, unsigned w)
{
    for (; *rcx; --*rcx, *rdi += w) memcpy(lift_wptr(c, *rdi, w), &rax, w);
}
static void lift_REPMOVS(cpu_t* c, uint64_t* rcx, uint64_t* rdi, uint64_t* rsi, unsigned w)
{
    for (; *rcx; --*rcx, *rdi += w, *rsi += w) memmove(lift_wptr(c, *rdi, w), lift_ptr(c, *rsi, w), w);
}
static inline uint64_t lift_CVTR32(double v) { return (v > -2147483649.0 && v < 2147483648.0) ? (uint64_t)(uint32_t)(int32_t)nearbyint(v) : 0x80000000ULL; }
static inline uint64_t lift_CVTR64(double v) { return (uint64_t)(int64_t)nearbyint(v); }

#ifdef LIFT_PROFILE
/* execution profile: instructions executed per translated function (analysis builds only) */
static struct { uint64_t fn, n; } g_prof[64];
static void lift_bb(uint64_t fn, unsigned n)
{
    for (int k = 0; k < 64; ++k) {
        if (g_prof[k].fn == fn || !g_prof[k].fn) { g_prof[k].fn = fn; g_prof[k].n += n; return; }
    }
}
#define LIFT_BB(fn, n) lift_bb(fn, n)
int cit_lifted_profile(uint64_t* fn, uint64_t* n, int reset)
{
    int k = 0;
    for (; k < 64 && g_prof[k].fn; ++k) { fn[k] = g_prof[k].fn; n[k] = g_prof[k].n; }
    if (reset) memset(g_prof, 0, sizeof g_prof);
    return k;
}
#endif
#include LIFT_GENERATED_INC

/* ---- instance API ------------------------------------------------------------------------------------------------ */
typedef struct cit_lifted {
    cpu_t cpu;
    uint8_t* mem;
} cit_lifted;

static const uint8_t* g_image;
static uint64_t g_image_size;

/* image: the bytes lift.py wrote (sections at their RVAs) */
int cit_lifted_set_image(const uint8_t* bytes, uint64_t n)
{
    if (n > LIFT_IMAGE_SIZE) return -1;
    uint8_t* copy = malloc(n);
    memcpy(copy, bytes, n);
    g_image = copy; g_image_size = n;
    return 0;
}

cit_lifted* cit_lifted_create(void)
{
    if (!g_image) return NULL;
    cit_lifted* m = calloc(1, sizeof *m);
    m->mem = calloc(1, LIFT_MEM_SIZE);
    memcpy(m->mem, g_image, g_image_size);
    m->cpu.M = m->mem;
    m->cpu.heap_next = LIFT_BASE + LIFT_IMAGE_SIZE;
    m->cpu.heap_end = LIFT_BASE + LIFT_IMAGE_SIZE + LIFT_HEAP_SIZE;
    return m;
}
void cit_lifted_destroy(cit_lifted* m) { if (m) { free(m->mem); free(m->cpu.wmask); free(m); } }

static void enter(cit_lifted* m)
{
    /* a fresh frame at the top of the stack: 16-byte aligned before the call pushes the return address, 32 bytes of
     * shadow space above it as the Windows x64 convention requires */
    m->cpu.r[4] = LIFT_BASE + LIFT_MEM_SIZE - 0x100 - 8;
}
void cit_lifted_initialize(cit_lifted* m) { m->cpu.guard = 0; enter(m); LIFT_INVOKE(f_1800096f0, &m->cpu); m->cpu.guard = 1; }
void cit_lifted_terminate_unguarded(cit_lifted* m) { m->cpu.guard = 0; }
void cit_lifted_terminate(cit_lifted* m) { m->cpu.guard = 0; enter(m); LIFT_INVOKE(f_18000e620, &m->cpu); }
/* step(out[12], in[11]): the buffers live in the emulated stack region above the frame */
void cit_lifted_step(cit_lifted* m, const double* in, double* out)
{
    const uint64_t a_in = LIFT_BASE + LIFT_MEM_SIZE - 0x100 + 0x20, a_out = a_in + 11 * 8 + 8;   /* 0x20 + 96 + 96 < 0x100 */
    memcpy(m->mem + (a_in - LIFT_BASE), in, 11 * sizeof(double));
    enter(m);
    m->cpu.r[1] = a_out; m->cpu.r[2] = a_in;
    LIFT_INVOKE(f_180003720, &m->cpu);
    memcpy(out, m->mem + (a_out - LIFT_BASE), 12 * sizeof(double));
}
void cit_lifted_run(cit_lifted* m, const double* in, int n_steps, double* out)
{
    for (int k = 0; k < n_steps; ++k) cit_lifted_step(m, in, out + 12 * k);
}
/* the 16 continuous states (RVAs established with the native binary, pe_citation.c) */
enum { CIT_RVA_X = 0x3c120, CIT_RVA_ENGINE = 0x3c198 };
void cit_lifted_get_state(cit_lifted* m, double* x12, double* eng4)
{
    if (x12) memcpy(x12, m->mem + CIT_RVA_X, 12 * sizeof(double));
    if (eng4) memcpy(eng4, m->mem + CIT_RVA_ENGINE, 4 * sizeof(double));
}
void cit_lifted_set_state(cit_lifted* m, const double* x12, const double* eng4)
{
    if (x12) memcpy(m->mem + CIT_RVA_X, x12, 12 * sizeof(double));
    if (eng4) memcpy(m->mem + CIT_RVA_ENGINE, eng4, 4 * sizeof(double));
}
void cit_lifted_onestep(cit_lifted* m, double* x, double* eng, const double* u, int n)
{
    double out[12];
    for (int i = 0; i < n; ++i) {
        cit_lifted_set_state(m, x + 12 * i, eng + 4 * i);
        cit_lifted_step(m, u + 11 * i, out);
        cit_lifted_get_state(m, x + 12 * i, eng + 4 * i);
    }
}
/* raw memory (tests, analysis): the whole emulated address space */
uint8_t* cit_lifted_memory(cit_lifted* m) { return m->mem; }
uint64_t cit_lifted_memory_size(void) { return LIFT_MEM_SIZE; }
uint64_t cit_lifted_heap_used(cit_lifted* m) { return m->cpu.heap_next - (LIFT_BASE + LIFT_IMAGE_SIZE); }
/* write tracking: one byte per byte of emulated memory, set when written */
uint8_t* cit_lifted_track_writes(cit_lifted* m, int on)
{
    if (on && !m->cpu.wmask) m->cpu.wmask = calloc(1, LIFT_MEM_SIZE);
    if (!on) { free(m->cpu.wmask); m->cpu.wmask = NULL; }
    return m->cpu.wmask;
}
