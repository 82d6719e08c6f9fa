/* rl4_citation_surrogate.h -- documented stand-in for the reference's nonlinear aircraft model.
 *
 * The reference's plant is `_citation.cp39-win_amd64.pyd` (envs/nonlinear/extended_input/, called at
 * envs/nonlinear/env.py:210,289-291 through envs/nonlinear/citation.py:62-69): a Simulink-Coder build
 * of the DASMAT Cessna Citation 500 model, Windows x64, CPython-3.9 ABI, NO source, fixed-step solver
 * string "ode5".  It cannot be IMPORTED here, but its model code runs in-process (oracle/pe_probe/): that
 * run is the truth this surrogate is calibrated against (tests/golden/citation_*.npz, profiles/citation_fidelity_r02.json).
 * The surrogate is NOT a restatement of the DASMAT model: parity with the reference plant is tolerance-level on the
 * identified envelope, not bit-level.  What IS reproduced exactly is its contract:
 *
 *   step(u[11]) -> x[12], one fixed step of dt seconds, process-global state in the reference /
 *   per-agent state here;
 *   x = [p q r V alpha beta phi theta psi h xe ye]      (envs/nonlinear/env.py:22-26, idhp_nonlin.py:52)
 *   u = [de da dr, trim de da dr, flap, gear, thr1, thr2, xcg shift]   (idhp_nonlin.py:51, env.py:130,143)
 *   trimmed straight and level at V = 90 m/s, h = 2000 m, alpha = theta = 0.0576 rad with
 *   de = -0.02855 rad, throttles 0.55 (idhp_nonlin.py:53-54).
 *
 * Model: rigid 6-DOF aeroplane, flat earth, ISA atmosphere, quasi-steady linear-in-derivatives aero
 * build-up with a quadratic drag polar and a smooth stall (lift saturation + post-stall drag / nose-down moment); stability and control derivatives of the Cessna Ce500 Citation
 * (TU Delft AE3202 Flight Dynamics lecture notes, Table D-1 -- the same table envs/linear/env.py:68-87
 * quotes), two fuselage-mounted turbofans lumped into one body-x thrust.  CL0, Cm0 and the static thrust
 * are solved so that the state above is an exact equilibrium of the model.  Integrators: classical RK4
 * (BASELINE.json "RK4 dt=0.01") and Dormand-Prince RK5 with fixed step (Simulink's "ode5").
 *
 * One definition, plain C, shared by the CUDA kernels (rl4afcs_b200/csrc/nl_kernels.cu) and the CPU
 * oracle's plant (oracle/nl_oracle.c): there is no reference to restate, so both sides must integrate
 * the SAME documented model.  It is written with IEEE basic operations only (+, -, *, /, sqrt, fma): sine / cosine
 * are the polynomial kernels below, the ISA power laws are binomial series whose coefficients sit in the parameter
 * block -- no libm / CUDA math-library call is left, so kernel and oracle produce the SAME BITS.
 */
#ifndef RL4_CITATION_SURROGATE_H
#define RL4_CITATION_SURROGATE_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RL4_HD __host__ __device__ __forceinline__
/* the derivative function is called 4-6 times per step: one out-of-line copy keeps the fused kernel inside the
 * instruction cache (fully inlined it was 370 KB of SASS and stalled on instruction fetch) */
#define RL4_HD_NOINLINE __host__ __device__ __forceinline__   /* out-of-line was tried: slower (x, u, k arrays forced to local memory) */
#define RL4_UNROLL _Pragma("unroll")
#else
#define RL4_HD static inline
#define RL4_HD_NOINLINE static inline
#define RL4_UNROLL
#endif

enum { RL4_CIT_NX = 12, RL4_CIT_NU = 11, RL4_CIT_NPOLY = 21 };
enum { RL4_CIT_P = 0, RL4_CIT_Q, RL4_CIT_R, RL4_CIT_V, RL4_CIT_ALPHA, RL4_CIT_BETA, RL4_CIT_PHI, RL4_CIT_THETA,
       RL4_CIT_PSI, RL4_CIT_H, RL4_CIT_XE, RL4_CIT_YE };
enum { RL4_CIT_INTEGRATOR_RK4 = 0, RL4_CIT_INTEGRATOR_ODE5 = 1 };

/* Longitudinal slopes identified against the reference's own plant binary (oracle/pe_probe/fit_surrogate.py --trajectories:
 * stage 1, least squares on the binary's one-step map over 20 000 states of the task envelope for the force slopes and
 * the thrust-speed slope; stage 2, trajectory match on four open-loop elevator manoeuvres for the three pitch-moment
 * slopes).  The AE3202 table values used in round 1 are quoted beside them: that table's aeroplane had 1.8x the elevator
 * power and a third of the pitch damping of the reference's. */
#define RL4_FIT_CLA  6.6284758345693255     /* AE3202: 5.16   */
#define RL4_FIT_CLQ  13.059503256533402     /* AE3202: 3.86   */
#define RL4_FIT_CLDE 0.41340524505785026    /* AE3202: 0.6238 */
#define RL4_FIT_CD0  0.015                  /* AE3202: 0.04   (lower bound of the fit) */
#define RL4_FIT_CDK  0.07643021209086866    /* AE3202: 0.052  */
#define RL4_FIT_CMA  -0.6634444312074602    /* AE3202: -0.43  */
#define RL4_FIT_CMQ  -25.0                  /* AE3202: -7.04  (bound of the fit: the binary's alpha-dot moments are lumped in) */
#define RL4_FIT_CMDE -1.3524530096077292    /* AE3202: -1.553 */
#define RL4_FIT_TV   -0.011411097863068305  /* relative thrust change per m/s of airspeed; round 1: 0 */
#define RL4_FIT_XCG  -0.8155089190491744    /* c.g. displacement per unit of input[10] (regression on the one-step map, r = -0.985); round 1: +1 */

typedef struct rl4_cit_params {
    /* mass and geometry */
    double m, S, c, b, Ixx, Iyy, Izz, Ixz, g;
    /* longitudinal derivatives (per rad, rates normalised with c/(2V)) */
    double CL0, CLa, CLq, CLde, CLflap, al_stall;
    double CD0, CDk, CDgear, CDflap, CDstall;
    double Cm0, Cma, Cmq, Cmde, Cmflap, Cmstall;
    /* lateral-directional derivatives (rates normalised with b/(2V)) */
    double CYb, CYp, CYr, CYda, CYdr;
    double Clb, Clp, Clr, Clda, Cldr;
    double Cnb, Cnp, Cnr, Cnda, Cndr;
    /* propulsion: T = Tstatic * (rho/rho0)^0.7 * (thr1 + thr2)/2 * (1 + TV (V - Vref)), along body x through the c.g.
     * (a turbofan's thrust falls with airspeed; TV is the relative slope per m/s) */
    double Tstatic, TV, Vref;
    /* input[10] (the reference's `shift_cg` fault sets it to -0.5, envs/nonlinear/env.py:141-143) -> c.g. displacement [m]:
     * sign and size identified against the binary (a NEGATIVE input pitches the reference's aircraft DOWN) */
    double xcg_gain;
    /* reciprocals used by rl4_cit_deriv, filled by rl4_cit_finalize() */
    double inv_m, inv_Iyy, inv_gam, inv_al_stall, inv_c, inv_b;
    /* ISA troposphere as binomial series in zeta = lapse h / T0, filled by rl4_cit_finalize():
     * rho / rho0 = (1 + zeta)^n_rho = sum rho_poly[k] zeta^k,  thrust lapse (rho / rho0)^0.7 = sum lapse_poly[k] zeta^k */
    double zeta_per_m;
    double rho_poly[RL4_CIT_NPOLY], lapse_poly[RL4_CIT_NPOLY];
} rl4_cit_params;

#define RL4_FMA(a, b, c) fma((a), (b), (c))
#ifndef RL4_DERIV_OUT_OF_LINE
#define RL4_DERIV_OUT_OF_LINE 0
#endif
#ifndef RL4_SINCOS_OUT_OF_LINE
#define RL4_SINCOS_OUT_OF_LINE 1   /* measured: fp64 kernel +27 %, mixed kernel -0.7 % against the inlined form */
#endif

/* derived constants; call after changing m, inertias, al_stall, c or b */
RL4_HD void rl4_cit_finalize(rl4_cit_params* P)
{
    const double T0 = 288.15, lapse = -0.0065, R = 287.05, g0 = 9.80665;
    const double n_rho = -(g0 / (lapse * R) + 1.0);       /* 4.2559 */
    const double n_lapse = 0.7 * n_rho;
    int k;
    P->inv_m = 1.0 / P->m; P->inv_Iyy = 1.0 / P->Iyy; P->inv_gam = 1.0 / (P->Ixx * P->Izz - P->Ixz * P->Ixz);
    P->inv_al_stall = 1.0 / P->al_stall; P->inv_c = 1.0 / P->c; P->inv_b = 1.0 / P->b;
    P->zeta_per_m = lapse / T0;
    P->rho_poly[0] = 1.0; P->lapse_poly[0] = 1.0;
    for (k = 1; k < RL4_CIT_NPOLY; ++k) {                  /* generalised binomial coefficients C(n, k) */
        P->rho_poly[k] = P->rho_poly[k - 1] * (n_rho - (double)(k - 1)) / (double)k;
        P->lapse_poly[k] = P->lapse_poly[k - 1] * (n_lapse - (double)(k - 1)) / (double)k;
    }
}

/* sin and cos from IEEE basic operations: Cody-Waite reduction by pi/2 with two FMAs (k = nearest integer to
 * a * 2/pi via the 1.5 * 2^52 shift), then the classical minimax kernels on [-pi/4, pi/4] (coefficients of the
 * fdlibm __kernel_sin / __kernel_cos polynomials), < 1.5 ulp for the angles of flight.  The same expression tree on
 * CUDA (-fmad=false, explicit FMAs) and on the host (-ffp-contract=off, fma()) gives the same bits. */
#define RL4_SC_LIST(X) \
    X(6.36619772367581382433e-01)  /* 0: 2/pi            */ \
    X(6755399441055744.0)          /* 1: 1.5 * 2^52      */ \
    X(1.57079632679489655800e+00)  /* 2: pi/2 high       */ \
    X(6.12323399573676603587e-17)  /* 3: pi/2 low        */ \
    X(-1.66666666666666324348e-01) /* 4..9: S1..S6       */ \
    X(8.33333333332248946124e-03)  \
    X(-1.98412698298579493134e-04) \
    X(2.75573137070700676789e-06)  \
    X(-2.50507602534068634195e-08) \
    X(1.58969099521155010221e-10)  \
    X(4.16666666666666019037e-02)  /* 10..15: C1..C6     */ \
    X(-1.38888888888741095749e-03) \
    X(2.48015872894767294178e-05)  \
    X(-2.75573143513906633035e-07) \
    X(2.08757232129817482790e-09)  \
    X(-1.13596475577881948265e-11)
#define RL4_SC_ELEM(v) v,
#if defined(__CUDACC__)
static __constant__ double rl4_sc_dev[16] = { RL4_SC_LIST(RL4_SC_ELEM) };   /* constant bank: an FMA operand, no moves */
#endif
static const double rl4_sc_host[16] = { RL4_SC_LIST(RL4_SC_ELEM) };
#if defined(__CUDA_ARCH__)
#define RL4_SC(i) rl4_sc_dev[i]
#define RL4_LOINT(t) __double2loint(t)
#else
#define RL4_SC(i) rl4_sc_host[i]
static inline int rl4_loint_host(double t) { int64_t b; memcpy(&b, &t, sizeof b); return (int)(uint32_t)(uint64_t)b; }
#define RL4_LOINT(t) rl4_loint_host(t)
#endif

RL4_HD void rl4_sincos(double a, double* s, double* c)
{
    const double t = RL4_FMA(a, RL4_SC(0), RL4_SC(1));
    const double kd = t - RL4_SC(1);
    const int k = RL4_LOINT(t);
    double r = RL4_FMA(-kd, RL4_SC(2), a);
    r = RL4_FMA(-kd, RL4_SC(3), r);
    {
        const double z = r * r;
        double ps = RL4_FMA(z, RL4_SC(9), RL4_SC(8)), pc = RL4_FMA(z, RL4_SC(15), RL4_SC(14));
        ps = RL4_FMA(z, ps, RL4_SC(7)); pc = RL4_FMA(z, pc, RL4_SC(13));
        ps = RL4_FMA(z, ps, RL4_SC(6)); pc = RL4_FMA(z, pc, RL4_SC(12));
        ps = RL4_FMA(z, ps, RL4_SC(5)); pc = RL4_FMA(z, pc, RL4_SC(11));
        ps = RL4_FMA(z, ps, RL4_SC(4)); pc = RL4_FMA(z, pc, RL4_SC(10));
        {
            const double sn = RL4_FMA(r * z, ps, r);
            const double cs = RL4_FMA(z * z, pc, RL4_FMA(z, -0.5, 1.0));
            const double S = (k & 1) ? cs : sn, C = (k & 1) ? sn : cs;
            *s = (k & 2) ? -S : S;
            *c = ((k + 1) & 2) ? -C : C;
        }
    }
}
/* Exact-zero shortcut (part of the definition, applied on both sides): sin(+-0) = +-0, cos(0) = 1.  In symmetric
 * flight (da = dr = 0) beta, phi and psi are identically zero, so three of the five evaluations per derivative are
 * skipped by a warp-uniform branch in the pitch-tracking task. */
#if defined(__CUDA_ARCH__) && RL4_SINCOS_OUT_OF_LINE
/* one out-of-line copy, values in registers (the derivative is inlined 4-6 times per step with five call sites each) */
static __device__ __noinline__ double2 rl4_sincos_ool(double a) { double2 r; rl4_sincos(a, &r.x, &r.y); return r; }
#define RL4_SINCOS(a, s, c) do { if ((a) == 0.0) { (s) = (a); (c) = 1.0; } else { const double2 sc_ = rl4_sincos_ool(a); (s) = sc_.x; (c) = sc_.y; } } while (0)
#else
#define RL4_SINCOS(a, s, c) do { if ((a) == 0.0) { (s) = (a); (c) = 1.0; } else { rl4_sincos((a), &(s), &(c)); } } while (0)
#endif

/* Air data frozen over one integration step (zero-order hold like the inputs): density and the thrust
 * lapse (rho/rho0)^0.7 are evaluated at the altitude at the START of the step; h changes by < 1 m per step.
 * The series are exact to 1e-16 relative from -2 km to 11 km (|zeta| < 0.25, 21 terms). */
typedef struct rl4_cit_air { double rho, thrust_lapse; } rl4_cit_air;
RL4_HD rl4_cit_air rl4_cit_airdata(const rl4_cit_params* P, double h)
{
    rl4_cit_air a;
    const double zeta = P->zeta_per_m * h;
    double pr = P->rho_poly[RL4_CIT_NPOLY - 1], pl = P->lapse_poly[RL4_CIT_NPOLY - 1];
    int k;
    RL4_UNROLL
    for (k = RL4_CIT_NPOLY - 2; k >= 0; --k) { pr = RL4_FMA(pr, zeta, P->rho_poly[k]); pl = RL4_FMA(pl, zeta, P->lapse_poly[k]); }
    a.rho = 1.225 * pr;
    a.thrust_lapse = pl;
    return a;
}
/* ISA troposphere density */
RL4_HD double rl4_cit_density(const rl4_cit_params* P, double h) { return rl4_cit_airdata(P, h).rho; }

/* xdot = f(x, u).  Written for the FP64 pipe: sums of products are explicit FMA chains (RL4_FMA = fma() on both
 * the CUDA and the host side), constant denominators are reciprocals precomputed in the parameter block, and the
 * four state-dependent reciprocals (1/V, 1/sqrt(1+(al/al_s)^2), 1/cos(theta), 1/cos(beta)) are formed once. */
RL4_HD_NOINLINE void rl4_cit_deriv(const rl4_cit_params* P, const rl4_cit_air air, const double* x, const double* u, double* dx)
{
    const double p = x[RL4_CIT_P], q = x[RL4_CIT_Q], r = x[RL4_CIT_R];
    const double V = x[RL4_CIT_V], al = x[RL4_CIT_ALPHA], be = x[RL4_CIT_BETA];
    const double phi = x[RL4_CIT_PHI], th = x[RL4_CIT_THETA], psi = x[RL4_CIT_PSI];
    const double de = u[0] + u[3], da = u[1] + u[4], dr = u[2] + u[5];
    const double flap = u[6], gear = u[7], thr = 0.5 * (u[8] + u[9]), dxcg = P->xcg_gain * u[10];

    double sa, ca, sb, cb, sphi, cphi, sth, cth, spsi, cpsi;
    RL4_SINCOS(al, sa, ca); RL4_SINCOS(be, sb, cb); RL4_SINCOS(phi, sphi, cphi); RL4_SINCOS(th, sth, cth); RL4_SINCOS(psi, spsi, cpsi);

    const double invV = 1.0 / V, inv_cth = 1.0 / cth, inv_cb = 1.0 / cb;
    const double qS = (0.5 * air.rho * P->S) * (V * V);
    const double ch = (0.5 * P->c) * invV, bh = (0.5 * P->b) * invV;
    const double qh = q * ch, ph = p * bh, rh = r * bh;

    /* aerodynamic coefficients; lift saturates smoothly beyond alpha_stall (al_e = al / sqrt(1 + (al/al_s)^2)),
     * the lost incidence (al - al_e) produces extra drag and a nose-down moment: a crude but bounded stall */
    const double an = al * P->inv_al_stall;
    const double al_e = al / sqrt(RL4_FMA(an, an, 1.0));
    const double al_x = al - al_e;
    const double CL = RL4_FMA(P->CLflap, flap, RL4_FMA(P->CLde, de, RL4_FMA(P->CLq, qh, RL4_FMA(P->CLa, al_e, P->CL0))));
    const double CD = RL4_FMA(P->CDstall * al_x, al_x, RL4_FMA(P->CDflap, flap, RL4_FMA(P->CDgear, gear, RL4_FMA(P->CDk * CL, CL, P->CD0))));
    const double CY = RL4_FMA(P->CYdr, dr, RL4_FMA(P->CYda, da, RL4_FMA(P->CYr, rh, RL4_FMA(P->CYp, ph, P->CYb * be))));
    const double CX = RL4_FMA(CL, sa, -CD * ca);          /* body axes */
    const double CZ = -RL4_FMA(CL, ca, CD * sa);
    /* a c.g. shift dxcg (m; the reference's `shift_cg` sets input[10] = -0.5) moves the moment reference:
     * dCm = CZ * dxcg / c, dCn = -CY * dxcg / b */
    const double Cm = RL4_FMA(CZ * dxcg, P->inv_c, RL4_FMA(P->Cmflap, flap, RL4_FMA(P->Cmde, de, RL4_FMA(P->Cmq, qh,
                      RL4_FMA(P->Cmstall, al_x, RL4_FMA(P->Cma, al, P->Cm0))))));
    const double Cl = RL4_FMA(P->Cldr, dr, RL4_FMA(P->Clda, da, RL4_FMA(P->Clr, rh, RL4_FMA(P->Clp, ph, P->Clb * be))));
    const double Cn = RL4_FMA(-CY * dxcg, P->inv_b, RL4_FMA(P->Cndr, dr, RL4_FMA(P->Cnda, da, RL4_FMA(P->Cnr, rh,
                      RL4_FMA(P->Cnp, ph, P->Cnb * be)))));

    const double T = (P->Tstatic * air.thrust_lapse * thr) * RL4_FMA(P->TV, V - P->Vref, 1.0);
    const double ax = RL4_FMA(qS, CX, T) * P->inv_m, ay = (qS * CY) * P->inv_m, az = (qS * CZ) * P->inv_m;   /* specific forces */
    const double L = (qS * P->b) * Cl, M = (qS * P->c) * Cm, N = (qS * P->b) * Cn;

    /* body-axis velocities and their rates */
    const double ub = V * ca * cb, vb = V * sb, wb = V * sa * cb;
    const double ud = RL4_FMA(r, vb, RL4_FMA(-q, wb, RL4_FMA(-P->g, sth, ax)));
    const double vd = RL4_FMA(p, wb, RL4_FMA(-r, ub, RL4_FMA(P->g * sphi, cth, ay)));
    const double wd = RL4_FMA(q, ub, RL4_FMA(-p, vb, RL4_FMA(P->g * cphi, cth, az)));
    const double Vd = RL4_FMA(ub, ud, RL4_FMA(vb, vd, wb * wd)) * invV;
    const double iVc = invV * inv_cb;                      /* 1 / sqrt(u^2 + w^2) */

    /* Euler's equations with Ixz */
    const double Lp = RL4_FMA(P->Ixz * p, q, RL4_FMA(-(P->Izz - P->Iyy) * q, r, L));
    const double Np = RL4_FMA(-P->Ixz * q, r, RL4_FMA(-(P->Iyy - P->Ixx) * p, q, N));
    const double qr = RL4_FMA(q, sphi, r * cphi);

    dx[RL4_CIT_P] = RL4_FMA(P->Izz, Lp, P->Ixz * Np) * P->inv_gam;
    dx[RL4_CIT_Q] = RL4_FMA(-P->Ixz, RL4_FMA(p, p, -r * r), RL4_FMA(-(P->Ixx - P->Izz) * p, r, M)) * P->inv_Iyy;
    dx[RL4_CIT_R] = RL4_FMA(P->Ixz, Lp, P->Ixx * Np) * P->inv_gam;
    dx[RL4_CIT_V] = Vd;
    dx[RL4_CIT_ALPHA] = RL4_FMA(ub, wd, -wb * ud) * (iVc * iVc);
    dx[RL4_CIT_BETA] = RL4_FMA(vd, V, -vb * Vd) * (invV * iVc);
    dx[RL4_CIT_PHI] = RL4_FMA(sth * inv_cth, qr, p);
    dx[RL4_CIT_THETA] = RL4_FMA(q, cphi, -r * sphi);
    dx[RL4_CIT_PSI] = qr * inv_cth;
    dx[RL4_CIT_H] = RL4_FMA(ub, sth, -RL4_FMA(vb * sphi, cth, (wb * cphi) * cth));
    {
        const double a1 = RL4_FMA(sphi * sth, cpsi, -cphi * spsi), a2 = RL4_FMA(cphi * sth, cpsi, sphi * spsi);
        const double b1 = RL4_FMA(sphi * sth, spsi, cphi * cpsi), b2 = RL4_FMA(cphi * sth, spsi, -sphi * cpsi);
        dx[RL4_CIT_XE] = RL4_FMA(ub * cth, cpsi, RL4_FMA(vb, a1, wb * a2));
        dx[RL4_CIT_YE] = RL4_FMA(ub * cth, spsi, RL4_FMA(vb, b1, wb * b2));
    }
}

#if defined(__CUDA_ARCH__) && RL4_DERIV_OUT_OF_LINE
/* ONE out-of-line copy of the derivative for the 4-6 stage evaluations of a step: scalars in, a 12-double struct out,
 * everything in registers (no pointers, so nothing is forced into local memory).  The fused kernel executes ~8 000
 * straight-line instructions per step, right at the instruction-cache capacity; sharing this code keeps it below. */
typedef struct rl4_cit_vec12 { double v[12]; } rl4_cit_vec12;
static __device__ __noinline__ rl4_cit_vec12 rl4_cit_deriv_ool(const rl4_cit_params* P, double rho, double lapse,
                                                                double x0, double x1, double x2, double x3, double x4, double x5,
                                                                double x6, double x7, double x8,
                                                                double u0, double u1, double u2, double u3, double u4, double u5,
                                                                double u6, double u7, double u8, double u9, double u10)
{
    const double x[12] = {x0, x1, x2, x3, x4, x5, x6, x7, x8, 0.0, 0.0, 0.0};
    const double u[11] = {u0, u1, u2, u3, u4, u5, u6, u7, u8, u9, u10};
    rl4_cit_air air; air.rho = rho; air.thrust_lapse = lapse;
    rl4_cit_vec12 d;
    rl4_cit_deriv(P, air, x, u, d.v);
    return d;
}
#define RL4_CIT_DERIV(P, air, x, u, dx) do { const rl4_cit_vec12 d_ = rl4_cit_deriv_ool((P), (air).rho, (air).thrust_lapse, \
    (x)[0], (x)[1], (x)[2], (x)[3], (x)[4], (x)[5], (x)[6], (x)[7], (x)[8], \
    (u)[0], (u)[1], (u)[2], (u)[3], (u)[4], (u)[5], (u)[6], (u)[7], (u)[8], (u)[9], (u)[10]); \
    RL4_UNROLL for (int i_ = 0; i_ < 12; ++i_) (dx)[i_] = d_.v[i_]; } while (0)
#else
#define RL4_CIT_DERIV(P, air, x, u, dx) rl4_cit_deriv((P), (air), (x), (u), (dx))
#endif

/* one fixed step, input held constant over the step (zero-order hold, like Simulink's fixed-step solvers);
 * stage combinations are explicit FMA chains (one rounding per term) */
RL4_HD void rl4_cit_step_rk4(const rl4_cit_params* P, double* x, const double* u, double dt)
{
    double k1[12], k2[12], k3[12], k4[12], y[12];
    int i;
    const rl4_cit_air air = rl4_cit_airdata(P, x[RL4_CIT_H]);
    const double hdt = 0.5 * dt, dt6 = dt / 6.0;
    RL4_CIT_DERIV(P, air, x, u, k1);
    RL4_UNROLL
    for (i = 0; i < 12; ++i) y[i] = RL4_FMA(hdt, k1[i], x[i]);
    RL4_CIT_DERIV(P, air, y, u, k2);
    RL4_UNROLL
    for (i = 0; i < 12; ++i) y[i] = RL4_FMA(hdt, k2[i], x[i]);
    RL4_CIT_DERIV(P, air, y, u, k3);
    RL4_UNROLL
    for (i = 0; i < 12; ++i) y[i] = RL4_FMA(dt, k3[i], x[i]);
    RL4_CIT_DERIV(P, air, y, u, k4);
    RL4_UNROLL
    for (i = 0; i < 12; ++i) x[i] = RL4_FMA(dt6, RL4_FMA(2.0, k2[i] + k3[i], k1[i] + k4[i]), x[i]);
}

/* Dormand-Prince 5(4) pair, 5th-order solution, fixed step: what Simulink calls ode5 */
RL4_HD void rl4_cit_step_ode5(const rl4_cit_params* P, double* x, const double* u, double dt)
{
    double k1[12], k2[12], k3[12], k4[12], k5[12], k6[12], y[12];
    int i;
    const rl4_cit_air air = rl4_cit_airdata(P, x[RL4_CIT_H]);
    const double dt5 = dt * (1.0 / 5.0);
    RL4_CIT_DERIV(P, air, x, u, k1);
    RL4_UNROLL
    for (i = 0; i < 12; ++i) y[i] = RL4_FMA(dt5, k1[i], x[i]);
    RL4_CIT_DERIV(P, air, y, u, k2);
    RL4_UNROLL
    for (i = 0; i < 12; ++i) y[i] = RL4_FMA(dt, RL4_FMA(9.0 / 40.0, k2[i], (3.0 / 40.0) * k1[i]), x[i]);
    RL4_CIT_DERIV(P, air, y, u, k3);
    RL4_UNROLL
    for (i = 0; i < 12; ++i)
        y[i] = RL4_FMA(dt, RL4_FMA(32.0 / 9.0, k3[i], RL4_FMA(-56.0 / 15.0, k2[i], (44.0 / 45.0) * k1[i])), x[i]);
    RL4_CIT_DERIV(P, air, y, u, k4);
    RL4_UNROLL
    for (i = 0; i < 12; ++i)
        y[i] = RL4_FMA(dt, RL4_FMA(-212.0 / 729.0, k4[i], RL4_FMA(64448.0 / 6561.0, k3[i],
                           RL4_FMA(-25360.0 / 2187.0, k2[i], (19372.0 / 6561.0) * k1[i]))), x[i]);
    RL4_CIT_DERIV(P, air, y, u, k5);
    RL4_UNROLL
    for (i = 0; i < 12; ++i)
        y[i] = RL4_FMA(dt, RL4_FMA(-5103.0 / 18656.0, k5[i], RL4_FMA(49.0 / 176.0, k4[i], RL4_FMA(46732.0 / 5247.0, k3[i],
                           RL4_FMA(-355.0 / 33.0, k2[i], (9017.0 / 3168.0) * k1[i])))), x[i]);
    RL4_CIT_DERIV(P, air, y, u, k6);
    RL4_UNROLL
    for (i = 0; i < 12; ++i)
        x[i] = RL4_FMA(dt, RL4_FMA(11.0 / 84.0, k6[i], RL4_FMA(-2187.0 / 6784.0, k5[i], RL4_FMA(125.0 / 192.0, k4[i],
                           RL4_FMA(500.0 / 1113.0, k3[i], (35.0 / 384.0) * k1[i])))), x[i]);
}

/* ---- symmetric flight --------------------------------------------------------------------------------------------------
 * With p = r = beta = phi = psi = 0 and da = dr = 0 (the pitch-tracking task: IDHPnonlin commands the elevator only,
 * objects.py:1448-1455) every lateral term of rl4_cit_deriv is an exact zero: products with a zero factor are +-0, an FMA
 * whose product is +-0 returns its addend, cos(0) = 1 and 1/1 = 1 are exact.  The functions below are rl4_cit_deriv /
 * rl4_cit_step_* with those terms dropped: for finite symmetric states they return the SAME values for
 * z = [q, V, alpha, theta, h, xe] (the lateral derivatives are +-0, so the lateral states stay zero) with ~40 % fewer
 * operations and half the stage storage.  tests/test_oracle_math.py compares them with the full model on the host,
 * the GPU parity tests compare the kernel (which takes this path when it applies) with the oracle (which never does). */
RL4_HD int rl4_cit_lon_in_range(const double* x)
{   /* far from overflow / division by zero: inside this box no intermediate of the step can become inf or nan, which is
     * what makes "zero times finite = zero" hold for every dropped lateral term (NaN compares false) */
    const double big = 1.0e30, tiny = 1.0e-30;
    return fabs(x[RL4_CIT_Q]) < big && fabs(x[RL4_CIT_V]) < big && fabs(x[RL4_CIT_V]) > tiny && fabs(x[RL4_CIT_ALPHA]) < big &&
           fabs(x[RL4_CIT_THETA]) < big && fabs(x[RL4_CIT_H]) < big && fabs(x[RL4_CIT_XE]) < big;
}
RL4_HD int rl4_cit_is_symmetric(const double* x, const double* u)
{
    return x[RL4_CIT_P] == 0.0 && x[RL4_CIT_R] == 0.0 && x[RL4_CIT_BETA] == 0.0 && x[RL4_CIT_PHI] == 0.0 && x[RL4_CIT_PSI] == 0.0 &&
           (u[1] + u[4]) == 0.0 && (u[2] + u[5]) == 0.0 && rl4_cit_lon_in_range(x);
}

/* z = [q, V, alpha, theta, h, xe];  c = [de, flap, gear, thr, dxcg] */
RL4_HD void rl4_cit_deriv_lon(const rl4_cit_params* P, const rl4_cit_air air, const double* z, const double* c, double* dz)
{
    const double q = z[0], V = z[1], al = z[2], th = z[3];
    const double de = c[0], flap = c[1], gear = c[2], thr = c[3], dxcg = P->xcg_gain * c[4];
    double sa, ca, sth, cth;
    /* The exact-zero shortcut of RL4_SINCOS as a select instead of a branch: rl4_sincos(+-0) = (+0, 1), the shortcut returns
     * (+-0, 1), so only the sign of a zero sine needs restoring.  The whole stage is then one branch-free block and the two
     * evaluations interleave (alpha and theta are never zero in flight). */
    rl4_sincos(al, &sa, &ca); rl4_sincos(th, &sth, &cth);
    sa = (al == 0.0) ? al : sa; sth = (th == 0.0) ? th : sth;
    const double invV = 1.0 / V;
    const double qS = (0.5 * air.rho * P->S) * (V * V);
    const double ch = (0.5 * P->c) * invV;
    const double qh = q * ch;
    const double an = al * P->inv_al_stall;
    const double al_e = al / sqrt(RL4_FMA(an, an, 1.0));
    const double al_x = al - al_e;
    const double CL = RL4_FMA(P->CLflap, flap, RL4_FMA(P->CLde, de, RL4_FMA(P->CLq, qh, RL4_FMA(P->CLa, al_e, P->CL0))));
    const double CD = RL4_FMA(P->CDstall * al_x, al_x, RL4_FMA(P->CDflap, flap, RL4_FMA(P->CDgear, gear, RL4_FMA(P->CDk * CL, CL, P->CD0))));
    const double CX = RL4_FMA(CL, sa, -CD * ca);
    const double CZ = -RL4_FMA(CL, ca, CD * sa);
    const double Cm = RL4_FMA(CZ * dxcg, P->inv_c, RL4_FMA(P->Cmflap, flap, RL4_FMA(P->Cmde, de, RL4_FMA(P->Cmq, qh,
                      RL4_FMA(P->Cmstall, al_x, RL4_FMA(P->Cma, al, P->Cm0))))));
    const double T = (P->Tstatic * air.thrust_lapse * thr) * RL4_FMA(P->TV, V - P->Vref, 1.0);
    const double ax = RL4_FMA(qS, CX, T) * P->inv_m, az = (qS * CZ) * P->inv_m;
    const double M = (qS * P->c) * Cm;
    const double ub = V * ca, wb = V * sa;                              /* cb = 1 */
    const double ud = RL4_FMA(-q, wb, RL4_FMA(-P->g, sth, ax));
    const double wd = RL4_FMA(q, ub, RL4_FMA(P->g, cth, az));           /* g * cos(phi) = g */
    const double Vd = RL4_FMA(ub, ud, wb * wd) * invV;
    dz[0] = M * P->inv_Iyy;
    dz[1] = Vd;
    dz[2] = RL4_FMA(ub, wd, -wb * ud) * (invV * invV);
    dz[3] = q;
    dz[4] = RL4_FMA(ub, sth, -(wb * cth));
    dz[5] = (ub * cth) + (wb * sth);
}

#define RL4_CIT_LON_PACK(x, u, z, c) do { \
    (z)[0] = (x)[RL4_CIT_Q]; (z)[1] = (x)[RL4_CIT_V]; (z)[2] = (x)[RL4_CIT_ALPHA]; (z)[3] = (x)[RL4_CIT_THETA]; \
    (z)[4] = (x)[RL4_CIT_H]; (z)[5] = (x)[RL4_CIT_XE]; \
    (c)[0] = (u)[0] + (u)[3]; (c)[1] = (u)[6]; (c)[2] = (u)[7]; (c)[3] = 0.5 * ((u)[8] + (u)[9]); (c)[4] = (u)[10]; } while (0)
#define RL4_CIT_LON_UNPACK(x, z) do { \
    (x)[RL4_CIT_Q] = (z)[0]; (x)[RL4_CIT_V] = (z)[1]; (x)[RL4_CIT_ALPHA] = (z)[2]; (x)[RL4_CIT_THETA] = (z)[3]; \
    (x)[RL4_CIT_H] = (z)[4]; (x)[RL4_CIT_XE] = (z)[5]; } while (0)

RL4_HD void rl4_cit_step_rk4_lon(const rl4_cit_params* P, double* x, const double* u, double dt)
{
    double z[6], c[5], k1[6], k2[6], k3[6], k4[6], y[6];
    int i;
    const rl4_cit_air air = rl4_cit_airdata(P, x[RL4_CIT_H]);
    const double hdt = 0.5 * dt, dt6 = dt / 6.0;
    RL4_CIT_LON_PACK(x, u, z, c);
    rl4_cit_deriv_lon(P, air, z, c, k1);
    RL4_UNROLL
    for (i = 0; i < 6; ++i) y[i] = RL4_FMA(hdt, k1[i], z[i]);
    rl4_cit_deriv_lon(P, air, y, c, k2);
    RL4_UNROLL
    for (i = 0; i < 6; ++i) y[i] = RL4_FMA(hdt, k2[i], z[i]);
    rl4_cit_deriv_lon(P, air, y, c, k3);
    RL4_UNROLL
    for (i = 0; i < 6; ++i) y[i] = RL4_FMA(dt, k3[i], z[i]);
    rl4_cit_deriv_lon(P, air, y, c, k4);
    RL4_UNROLL
    for (i = 0; i < 6; ++i) z[i] = RL4_FMA(dt6, RL4_FMA(2.0, k2[i] + k3[i], k1[i] + k4[i]), z[i]);
    RL4_CIT_LON_UNPACK(x, z);
}

RL4_HD void rl4_cit_step_ode5_lon(const rl4_cit_params* P, double* x, const double* u, double dt)
{
    double z[6], c[5], k1[6], k2[6], k3[6], k4[6], k5[6], k6[6], y[6];
    int i;
    const rl4_cit_air air = rl4_cit_airdata(P, x[RL4_CIT_H]);
    const double dt5 = dt * (1.0 / 5.0);
    RL4_CIT_LON_PACK(x, u, z, c);
    rl4_cit_deriv_lon(P, air, z, c, k1);
    RL4_UNROLL
    for (i = 0; i < 6; ++i) y[i] = RL4_FMA(dt5, k1[i], z[i]);
    rl4_cit_deriv_lon(P, air, y, c, k2);
    RL4_UNROLL
    for (i = 0; i < 6; ++i) y[i] = RL4_FMA(dt, RL4_FMA(9.0 / 40.0, k2[i], (3.0 / 40.0) * k1[i]), z[i]);
    rl4_cit_deriv_lon(P, air, y, c, k3);
    RL4_UNROLL
    for (i = 0; i < 6; ++i)
        y[i] = RL4_FMA(dt, RL4_FMA(32.0 / 9.0, k3[i], RL4_FMA(-56.0 / 15.0, k2[i], (44.0 / 45.0) * k1[i])), z[i]);
    rl4_cit_deriv_lon(P, air, y, c, k4);
    RL4_UNROLL
    for (i = 0; i < 6; ++i)
        y[i] = RL4_FMA(dt, RL4_FMA(-212.0 / 729.0, k4[i], RL4_FMA(64448.0 / 6561.0, k3[i],
                           RL4_FMA(-25360.0 / 2187.0, k2[i], (19372.0 / 6561.0) * k1[i]))), z[i]);
    rl4_cit_deriv_lon(P, air, y, c, k5);
    RL4_UNROLL
    for (i = 0; i < 6; ++i)
        y[i] = RL4_FMA(dt, RL4_FMA(-5103.0 / 18656.0, k5[i], RL4_FMA(49.0 / 176.0, k4[i], RL4_FMA(46732.0 / 5247.0, k3[i],
                           RL4_FMA(-355.0 / 33.0, k2[i], (9017.0 / 3168.0) * k1[i])))), z[i]);
    rl4_cit_deriv_lon(P, air, y, c, k6);
    RL4_UNROLL
    for (i = 0; i < 6; ++i)
        z[i] = RL4_FMA(dt, RL4_FMA(11.0 / 84.0, k6[i], RL4_FMA(-2187.0 / 6784.0, k5[i], RL4_FMA(125.0 / 192.0, k4[i],
                           RL4_FMA(500.0 / 1113.0, k3[i], (35.0 / 384.0) * k1[i])))), z[i]);
    RL4_CIT_LON_UNPACK(x, z);
}

/* One plant step that takes the symmetric-flight form when it provably equals the full model: symmetric, in-range state
 * before the step AND an in-range result (a step that blows up is redone with the full equations, whose lateral terms
 * then turn into NaN exactly as the oracle's do). */
RL4_HD void rl4_cit_step_auto(const rl4_cit_params* P, double* x, const double* u, double dt, int integrator)
{
    if (rl4_cit_is_symmetric(x, u)) {
        const double s0 = x[RL4_CIT_Q], s1 = x[RL4_CIT_V], s2 = x[RL4_CIT_ALPHA], s3 = x[RL4_CIT_THETA], s4 = x[RL4_CIT_H], s5 = x[RL4_CIT_XE];
        if (integrator == RL4_CIT_INTEGRATOR_RK4) rl4_cit_step_rk4_lon(P, x, u, dt); else rl4_cit_step_ode5_lon(P, x, u, dt);
        if (rl4_cit_lon_in_range(x)) return;
        x[RL4_CIT_Q] = s0; x[RL4_CIT_V] = s1; x[RL4_CIT_ALPHA] = s2; x[RL4_CIT_THETA] = s3; x[RL4_CIT_H] = s4; x[RL4_CIT_XE] = s5;
    }
    if (integrator == RL4_CIT_INTEGRATOR_RK4) rl4_cit_step_rk4(P, x, u, dt); else rl4_cit_step_ode5(P, x, u, dt);
}

/* CL0, Cm0 and the static thrust for which (V, h, alpha = theta, de, throttle) = (90, 2000, 0.0576, -0.02855, 0.55) -- the
 * trim point of idhp_nonlin.py:53-54 -- is an exact equilibrium of the model with the slopes currently in *P:
 * T cos(al) = D, L + T sin(al) = W with D = qS (CD0 + k CL^2 + ...), L = qS CL:
 * k/qS (W - T sa)^2 + qS CD0 - T ca = 0  ->  quadratic in T, smaller root. */
RL4_HD void rl4_cit_solve_trim(rl4_cit_params* P)
{
    const double V = 90.0, h = 2000.0, al = 0.0576, de = -0.02855, thr = 0.55;
    double rho, lapse, qS, sa, ca, W, T, CLt, A, B, C, al_e, al_x;
    rl4_cit_finalize(P);                                   /* atmosphere tables are needed by the trim solve */
    {
        const rl4_cit_air air = rl4_cit_airdata(P, h);
        rho = air.rho; lapse = air.thrust_lapse;
    }
    qS = 0.5 * rho * V * V * P->S; rl4_sincos(al, &sa, &ca); W = P->m * P->g;
    al_e = al / sqrt(1.0 + (al / P->al_stall) * (al / P->al_stall)); al_x = al - al_e;
    A = P->CDk / qS * sa * sa; B = -(2.0 * P->CDk / qS * W * sa + ca);
    C = P->CDk / qS * W * W + qS * (P->CD0 + P->CDstall * al_x * al_x);
    T = (-B - sqrt(B * B - 4.0 * A * C)) / (2.0 * A);
    CLt = (W - T * sa) / qS;
    P->CL0 = CLt - P->CLa * al_e - P->CLde * de;
    P->Cm0 = -(P->Cma * al + P->Cmstall * al_x + P->Cmde * de);
    P->Vref = V;
    P->Tstatic = T / (lapse * thr);                        /* the speed factor 1 + TV (V - Vref) is one at the trim speed */
    rl4_cit_finalize(P);
}

/* Default parameter set.  Geometry, mass and the lateral-directional derivatives are the Cessna Ce500 Citation table of the
 * TU Delft AE3202 lecture notes (the table envs/linear/env.py:68-87 quotes).  The LONGITUDINAL slopes and the thrust-speed
 * slope are identified against the one-step map of the reference's own plant binary, run in-process by oracle/pe_probe
 * (fit_surrogate.py: 20 000 samples over the envelope of the pitch-tracking task, least squares inside a physically
 * plausible box); CL0, Cm0 and the static thrust then follow from the reference's trim point. */
RL4_HD void rl4_cit_default_params(rl4_cit_params* P)
{
    P->m = 4547.8; P->S = 24.2; P->c = 2.022; P->b = 13.36; P->g = 9.80665;
    P->Ixx = P->m * P->b * P->b * 0.012; P->Izz = P->m * P->b * P->b * 0.037; P->Ixz = P->m * P->b * P->b * 0.002;
    P->Iyy = P->m * P->c * P->c * 0.980;       /* K_Y^2 = Iyy / (m c^2) */
    P->CLa = RL4_FIT_CLA; P->CLq = RL4_FIT_CLQ; P->CLde = RL4_FIT_CLDE; P->CLflap = 0.6; P->al_stall = 0.28;
    P->CD0 = RL4_FIT_CD0; P->CDk = RL4_FIT_CDK; P->CDgear = 0.02; P->CDflap = 0.04; P->CDstall = 8.0;
    P->Cma = RL4_FIT_CMA; P->Cmq = RL4_FIT_CMQ; P->Cmde = RL4_FIT_CMDE; P->Cmflap = -0.05; P->Cmstall = -6.0;
    P->CYb = -0.9896; P->CYp = -0.087; P->CYr = 0.43; P->CYda = 0.0; P->CYdr = 0.3037;
    P->Clb = -0.0772; P->Clp = -0.3444; P->Clr = 0.28; P->Clda = -0.2349; P->Cldr = 0.0286;
    P->Cnb = 0.1638; P->Cnp = -0.0108; P->Cnr = -0.193; P->Cnda = 0.0286; P->Cndr = -0.1261;
    P->TV = RL4_FIT_TV; P->xcg_gain = RL4_FIT_XCG;
    rl4_cit_solve_trim(P);
}

#endif
