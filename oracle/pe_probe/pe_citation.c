/* TEST INFRASTRUCTURE ONLY -- in-process loader for the reference's nonlinear plant binary.
 *
 * The reference's aircraft model is `envs/nonlinear/<variant>/_citation.cp39-win_amd64.pyd` (called at
 * envs/nonlinear/env.py:210,288-291 through envs/nonlinear/citation.py:62-69): a Windows x86-64 PE DLL holding a
 * Simulink-Coder build of the DASMAT Citation model behind a SWIG/CPython-3.9 wrapper, no source.  It cannot be
 * imported on Linux / CPython 3.12 -- but the MODEL CODE inside it is plain x86-64 machine code that needs only eight
 * libm functions and memcpy / memset.  This file maps the image (sections at their RVAs, base relocations applied when
 * the preferred base is taken), binds the C-runtime imports to glibc through ms_abi thunks, leaves every python39.dll /
 * KERNEL32 import pointing at a trap, and calls the three model entry points directly, bypassing the SWIG layer:
 *
 *     RVA 0x96f0   void initialize(void)                      <- _wrap_initialize  (RVA 0x101a0: `call 0x1800096f0`)
 *     RVA 0x3720   void step(double out[12], const double in[11])   <- _wrap_step  (RVA 0x102c0: rcx = data of the new
 *                                                                 12-element array, rdx = data of the input array,
 *                                                                 `call 0x180003720` at RVA 0x1063d)
 *     RVA 0xe620   void terminate(void)                       <- _wrap_terminate   (RVA 0x10230: `call 0x18000e620`)
 *
 * (addresses read from `objdump -d -M intel` of the PyMethodDef table at RVA 0x2e0f0; oracle/pe_probe/README.md has the
 * recipe; the four variants of the binary share the layout, pe_citation_open() checks the call sites before trusting it).
 * The DLL entry point (CRT start-up) is never run: the model code uses no CRT state.
 *
 * Build: oracle/pe_probe/Makefile -> oracle/_ref/libpe_citation.so.  Only the build container has /root/reference.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#define MSABI __attribute__((ms_abi))

typedef struct {
    uint8_t* base;
    size_t size;
    void (MSABI *initialize)(void);
    void (MSABI *step)(double*, const double*);
    void (MSABI *terminate)(void);
    char trapped[128];
} pe_image;

static pe_image g_img;
static char g_err[256];

/* ---- imports the model code may call: C runtime through ms_abi thunks ---- */
static double MSABI w_cos(double x) { return cos(x); }
static double MSABI w_sin(double x) { return sin(x); }
static double MSABI w_tan(double x) { return tan(x); }
static double MSABI w_exp(double x) { return exp(x); }
static double MSABI w_floor(double x) { return floor(x); }
static double MSABI w_log10(double x) { return log10(x); }
static double MSABI w_sqrt(double x) { return sqrt(x); }
static double MSABI w_pow(double x, double y) { return pow(x, y); }
static void* MSABI w_memcpy(void* d, const void* s, size_t n) { return memcpy(d, s, n); }
static void* MSABI w_memset(void* d, int c, size_t n) { return memset(d, c, n); }
static void* MSABI w_malloc(size_t n) { return malloc(n); }
static void MSABI w_free(void* p) { free(p); }
static int MSABI w_strcmp(const char* a, const char* b) { return strcmp(a, b); }
static int MSABI w_strncmp(const char* a, const char* b, size_t n) { return strncmp(a, b, n); }
static char* MSABI w_strstr(const char* a, const char* b) { return strstr(a, b); }
/* anything else (python39.dll, KERNEL32, CRT start-up) must never be reached from the model code */
static void MSABI w_trap(void)
{
    fprintf(stderr, "pe_citation: the model code called an import that is not bound (python / kernel32 / CRT start-up)\n");
    abort();
}
static uint8_t g_dummy_data[4096];      /* data imports (PyExc_*, _Py_NoneStruct ...) point here; never dereferenced by the model */

static const struct { const char* name; void* fn; } kImports[] = {
    {"cos", (void*)w_cos}, {"sin", (void*)w_sin}, {"tan", (void*)w_tan}, {"exp", (void*)w_exp}, {"floor", (void*)w_floor},
    {"log10", (void*)w_log10}, {"sqrt", (void*)w_sqrt}, {"pow", (void*)w_pow}, {"memcpy", (void*)w_memcpy},
    {"memset", (void*)w_memset}, {"malloc", (void*)w_malloc}, {"free", (void*)w_free}, {"strcmp", (void*)w_strcmp},
    {"strncmp", (void*)w_strncmp}, {"strstr", (void*)w_strstr},
};

static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
static uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

const char* pe_citation_error(void) { return g_err; }

/* Maps `path`; returns 0 on success.  n_bound / n_trapped report how many imports went to glibc / to the trap. */
int pe_citation_open(const char* path, int* n_bound, int* n_trapped)
{
    FILE* f = fopen(path, "rb");
    if (!f) { snprintf(g_err, sizeof g_err, "cannot open %s", path); return -1; }
    fseek(f, 0, SEEK_END);
    long fsz = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t* file = malloc((size_t)fsz);
    if (fread(file, 1, (size_t)fsz, f) != (size_t)fsz) { fclose(f); snprintf(g_err, sizeof g_err, "short read"); return -1; }
    fclose(f);
    if (fsz < 0x200 || file[0] != 'M' || file[1] != 'Z') { snprintf(g_err, sizeof g_err, "not a PE file"); return -1; }
    const uint8_t* nt = file + rd32(file + 0x3c);
    if (rd32(nt) != 0x4550 || rd16(nt + 4) != 0x8664) { snprintf(g_err, sizeof g_err, "not a PE32+ x86-64 image"); return -1; }
    const int nsec = rd16(nt + 6);
    const int optsz = rd16(nt + 20);
    const uint8_t* opt = nt + 24;
    if (rd16(opt) != 0x20b) { snprintf(g_err, sizeof g_err, "not PE32+"); return -1; }
    const uint64_t pref = rd64(opt + 24);
    const uint32_t image_size = rd32(opt + 56), hdr_size = rd32(opt + 60);
    const uint8_t* dirs = opt + 112;                 /* data directories: [1] import, [5] base relocation */
    uint8_t* base = mmap((void*)pref, image_size, PROT_READ | PROT_WRITE | PROT_EXEC,
                         MAP_PRIVATE | MAP_ANONYMOUS | MAP_FIXED_NOREPLACE, -1, 0);
    if (base == MAP_FAILED)
        base = mmap(NULL, image_size, PROT_READ | PROT_WRITE | PROT_EXEC, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (base == MAP_FAILED) { snprintf(g_err, sizeof g_err, "mmap failed"); return -1; }
    memcpy(base, file, hdr_size);
    const uint8_t* sec = opt + optsz;
    for (int i = 0; i < nsec; ++i, sec += 40) {
        const uint32_t vsz = rd32(sec + 8), va = rd32(sec + 12), rsz = rd32(sec + 16), ro = rd32(sec + 20);
        memcpy(base + va, file + ro, rsz < vsz ? rsz : vsz);
    }
    /* base relocations (only when the preferred base was not available) */
    const int64_t delta = (int64_t)((uint64_t)(uintptr_t)base - pref);
    if (delta) {
        const uint32_t rva = rd32(dirs + 5 * 8), sz = rd32(dirs + 5 * 8 + 4);
        for (uint32_t o = 0; o + 8 <= sz;) {
            const uint32_t page = rd32(base + rva + o), blk = rd32(base + rva + o + 4);
            if (blk < 8) break;
            for (uint32_t k = 8; k + 2 <= blk; k += 2) {
                const uint16_t e = rd16(base + rva + o + k);
                if ((e >> 12) == 10) { uint64_t v = rd64(base + page + (e & 0xfff)); v += (uint64_t)delta; memcpy(base + page + (e & 0xfff), &v, 8); }
            }
            o += blk;
        }
    }
    /* imports */
    int nb = 0, nt_ = 0;
    const uint32_t irva = rd32(dirs + 1 * 8);
    for (const uint8_t* d = base + irva; rd32(d + 12); d += 20) {
        const uint32_t oft = rd32(d) ? rd32(d) : rd32(d + 16), ft = rd32(d + 16);
        for (uint32_t k = 0;; k += 8) {
            const uint64_t ent = rd64(base + oft + k);
            if (!ent) break;
            void* target = (void*)w_trap;
            if (!(ent >> 63)) {
                const char* name = (const char*)(base + (uint32_t)ent + 2);
                int found = 0;
                for (size_t j = 0; j < sizeof kImports / sizeof kImports[0]; ++j)
                    if (!strcmp(name, kImports[j].name)) { target = kImports[j].fn; found = 1; break; }
                if (!found && (!strncmp(name, "PyExc_", 6) || !strncmp(name, "_Py_", 4) || !strcmp(name, "PyCapsule_Type"))) target = g_dummy_data;
                if (found) ++nb; else ++nt_;
            } else ++nt_;
            memcpy(base + ft + k, &target, 8);
        }
    }
    /* the model entry points, checked against the call sites in the SWIG wrappers (E8 rel32) */
    const struct { uint32_t site, target; } calls[3] = {{0x1021a, 0x96f0}, {0x1063d, 0x3720}, {0x102aa, 0xe620}};
    for (int i = 0; i < 3; ++i) {
        const uint8_t* s = base + calls[i].site;
        int32_t rel; memcpy(&rel, s + 1, 4);
        if (s[0] != 0xE8 || (uint32_t)(calls[i].site + 5 + rel) != calls[i].target) {
            snprintf(g_err, sizeof g_err, "call site 0x%x does not call RVA 0x%x: another build of the binary", calls[i].site, calls[i].target);
            munmap(base, image_size);
            return -2;
        }
    }
    g_img.base = base; g_img.size = image_size;
    g_img.initialize = (void (MSABI*)(void))(base + 0x96f0);
    g_img.step = (void (MSABI*)(double*, const double*))(base + 0x3720);
    g_img.terminate = (void (MSABI*)(void))(base + 0xe620);
    if (n_bound) *n_bound = nb;
    if (n_trapped) *n_trapped = nt_;
    free(file);
    return 0;
}

int pe_citation_initialize(void) { if (!g_img.base) return -1; g_img.initialize(); return 0; }
int pe_citation_terminate(void) { if (!g_img.base) return -1; g_img.terminate(); return 0; }
/* citation.step(cmd): in[11] = [de da dr, trim de da dr, flap, gear, thr1, thr2, xcg] -> out[12] (envs/nonlinear/env.py:22-26) */
int pe_citation_step(const double* in, double* out) { if (!g_img.base) return -1; g_img.step(out, in); return 0; }
/* n_steps with the same input; out [n_steps][12] */
int pe_citation_run(const double* in, int n_steps, double* out)
{
    if (!g_img.base) return -1;
    for (int k = 0; k < n_steps; ++k) g_img.step(out + 12 * k, in);
    return 0;
}
uint64_t pe_citation_base(void) { return (uint64_t)(uintptr_t)g_img.base; }

/* The model's continuous states live in .data: the 12 airframe states [p q r V alpha beta phi theta psi h xe ye] at RVA
 * 0x3c120 and four engine states (two per engine) at RVA 0x3c198 -- found by perturbing every double of .data that
 * changes from step to step and watching the outputs (oracle/pe_probe/README.md).  The first step() after initialize()
 * emits the initial condition; from then on step() integrates from whatever these 16 doubles hold, which is what makes
 * the one-step map x_next = F(x, engine, u) of the real plant observable from arbitrary states. */
enum { PE_CIT_RVA_X = 0x3c120, PE_CIT_RVA_ENGINE = 0x3c198 };
int pe_citation_get_state(double* x12, double* eng4)
{
    if (!g_img.base) return -1;
    if (x12) memcpy(x12, g_img.base + PE_CIT_RVA_X, 12 * sizeof(double));
    if (eng4) memcpy(eng4, g_img.base + PE_CIT_RVA_ENGINE, 4 * sizeof(double));
    return 0;
}
int pe_citation_set_state(const double* x12, const double* eng4)
{
    if (!g_img.base) return -1;
    if (x12) memcpy(g_img.base + PE_CIT_RVA_X, x12, 12 * sizeof(double));
    if (eng4) memcpy(g_img.base + PE_CIT_RVA_ENGINE, eng4, 4 * sizeof(double));
    return 0;
}
/* x_next = F(x, engine, u) for n samples: x [n][12], eng [n][4] (in: state before, out: state after), u [n][11] */
int pe_citation_onestep(double* x, double* eng, const double* u, int n)
{
    if (!g_img.base) return -1;
    double out[12];
    for (int i = 0; i < n; ++i) {
        pe_citation_set_state(x + 12 * i, eng + 4 * i);
        g_img.step(out, u + 11 * i);
        pe_citation_get_state(x + 12 * i, eng + 4 * i);
    }
    return 0;
}
