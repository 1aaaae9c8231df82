/**
 * iso2d_device.cuh -- point-wise device physics of the isothermal-2D `binary` update.
 *
 * fp64 throughout.  Each function names the reference routine it replaces
 * (paths relative to Mara3 src/).  The arithmetic is algebraically the
 * reference's but strength-reduced for the B200 fp64 pipe (the stage kernel is
 * fp64-issue bound, not HBM bound): reciprocals / reciprocal square roots are a
 * MUFU seed plus one cubically convergent correction (relative error < 1e-15),
 * pow(x, 1.5) and pow(x, 0.5) become products of rsqrt, and exp() in the sink
 * term is skipped where it underflows any possible contribution.  The allowed
 * mismatch against the reference is 1e-12 relative per cell per step; the parity
 * tests in tests/ measure it.
 */
#pragma once
#ifdef M3B_HOST_EMULATION
// tests/ only (tests/device_math_host.cpp): the point-wise functions below compiled by the host compiler so that their
// algebra can be checked against the oracle without a GPU.  Hardware seeds become exact library calls; nothing else changes.
#include <cmath>
#include <cstring>
#define __device__
#define __forceinline__ inline
#define __noinline__
static inline int __double2hiint(double x) { long long b; std::memcpy(&b, &x, 8); return int(b >> 32); }
static inline int __double2loint(double x) { long long b; std::memcpy(&b, &x, 8); return int(b & 0xffffffffll); }
static inline double __hiloint2double(int hi, int lo) { long long b = (static_cast<long long>(hi) << 32) | static_cast<unsigned int>(lo); double x; std::memcpy(&x, &b, 8); return x; }
using std::fma; using std::fabs; using std::sqrt; using std::exp; using std::tanh;
static inline int __any_sync(unsigned, int p) { return p; }
static inline double __shfl_xor_sync(unsigned, double, int) { return 0.0; }     // a "warp" of one lane: the others hold nothing
static const struct { unsigned x; } threadIdx = {0};
#else
#include <cuda_runtime.h>
#endif
#include "iso2d_sums.hpp"

namespace m3b { namespace dev {

/** Per-run physical constants (solver_data_t scalars, subprog_binary.hpp:76-97). */
struct model_t
{
    double softening_radius2;   // rs^2
    double sink_rate;
    double sink_inv_2s2;        // 1 / (2 sink_radius^2)
    double inv_mach2;           // 1 / M^2
    double inv_mach;
    double alpha, nu, alpha_cutoff_radius;
    double density_floor;
    int axisymmetric_cs2;
    // conserve_linear_p = 0: the state is conserved_q = (sigma, Sr, Lz) (advance_q, scheme.cpp:906-1020)
    int qmode;
    double domain_radius;       // no Lz flux through faces at x, y = +- domain_radius (scheme.cpp:209-210)
    double gst_suppr_radius2;   // ramp of the geometric source term (scheme.cpp:424, 439-444)
};

/** Per-stage inputs (body positions at the stage time; scheme.cpp:814). */
struct stage_t
{
    double time;        // solution.time of the stage input
    double dt;
    double theta;       // plm_theta, or 0 in safe mode (scheme.cpp:792)
    double x1, y1, m1;  // body 1 position, mass
    double x2, y2, m2;  // body 2
    double vx1, vy1, vx2, vy2;  // body velocities (work integral, scheme.cpp:363-374)
    double rk_b0;       // if combine: out = Un * b0 + updated * (1 - b0)   (subprog_binary.cpp:272-275)
    int combine;
    int compute_dt;     // also reduce spacing / max wavespeed of the OUTPUT state (scheme.cpp:1107-1126)
};

struct prim_t { double s, vx, vy; };

// ---------------------------------------------------------------------------
// fast reciprocal / rsqrt: hardware seed (~2^-20) + one third-order step
// ---------------------------------------------------------------------------
__device__ __forceinline__ double fast_rcp(double x)
{
#ifdef M3B_HOST_EMULATION
    return 1.0 / x;
#endif
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    double t = fma(e, e, e);
    return fma(y, t, y);
}

__device__ __forceinline__ double fast_rsqrt(double x)
{
#ifdef M3B_HOST_EMULATION
    return 1.0 / std::sqrt(x);
#endif
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double t = x * y;
    double e = fma(-t, y, 1.0);                  // e = 1 - x y^2
    double p = fma(0.375, e, 0.5);
    return fma(y, p * e, y);                     // y (1 + e/2 + 3 e^2 / 8)
}

/** sqrt(x) from the same seed: g = x y is already the root to seed accuracy, and the residual 1 - g y = 1 - x y^2 corrects it
 *  to third order exactly as it corrects y -- one multiplication less than x * fast_rsqrt(x). */
__device__ __forceinline__ double fast_sqrt(double x)
{
#ifdef M3B_HOST_EMULATION
    return std::sqrt(x);
#endif
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y;
    double e = fma(-g, y, 1.0);
    double p = fma(0.375, e, 0.5);
    return fma(g, p * e, g);
}

__device__ __forceinline__ prim_t cons_to_prim(double s, double px, double py)
{
    // iso2d::recover_primitive (physics_iso2d.hpp:351-362): (sigma, px / sigma, py / sigma)
    double inv = fast_rcp(s);
    return {s, px * inv, py * inv};
}

/** min / max by compare-and-select: no NaN canonicalisation (DSETP + 2 SEL instead of ~6 instructions). */
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }

/** max(0, x) and min(0, x) on the bit pattern (a shift and two ANDs on the integer pipes instead of DSETP + 2 SEL:
 *  the stage kernels are bound by the fp64 pipe, so every compare moved off it is two issue cycles back). */
__device__ __forceinline__ double dmax0(double x)
{
    const int hi = __double2hiint(x), m = ~(hi >> 31);
    return __hiloint2double(hi & m, __double2loint(x) & m);
}
__device__ __forceinline__ double dmin0(double x)
{
    const int hi = __double2hiint(x), m = hi >> 31;
    return __hiloint2double(hi & m, __double2loint(x) & m);
}

#ifndef M3B_HOST_EMULATION
/** Eight sums over the 32 lanes of a warp by recursive halving: at distances 16, 8, 4 a lane keeps half of its values and trades
 *  the other half, then two butterfly steps finish the one value that is left.  Lane 4 j returns the sum of v8[j]. */
__device__ __forceinline__ double warp_sum8(const double v8[8], int lane)
{
    double v4[4], v2[2], v1;
    const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
    #pragma unroll
    for (int q = 0; q < 4; ++q)
    {
        const double lo = v8[q], hi = v8[4 + q];
        v4[q] = (b16 ? hi : lo) + __shfl_xor_sync(0xffffffffu, b16 ? lo : hi, 16);
    }
    #pragma unroll
    for (int q = 0; q < 2; ++q) v2[q] = (b8 ? v4[2 + q] : v4[q]) + __shfl_xor_sync(0xffffffffu, b8 ? v4[q] : v4[2 + q], 8);
    v1 = (b4 ? v2[1] : v2[0]) + __shfl_xor_sync(0xffffffffu, b4 ? v2[0] : v2[1], 4);
    v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
    v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
    return v1;
}
#endif

/**
 * Un-divided PLM difference: mara::plm_gradient (math_interpolation.hpp:85-94),
 *   0.25 |sgn a + sgn b| (sgn a + sgn c) min(|a|, |b|, |c|),  a = theta dl, b = (dl + dr) / 2, c = theta dr
 * with dl = y0 - yl, dr = yr - y0.  b carries the common sign whenever dl and dr agree in sign and
 * the result is zero when they do not, so it equals sgn(dl) min(theta min(|dl|, |dr|), (|dl| + |dr|) / 2)
 * when the sign bits of dl and dr agree and 0 otherwise.
 */
__device__ __forceinline__ double plm_from_differences(double dl, double dr, double theta)
{
    // signed throughout: t = theta * (the smaller of dl, dr in magnitude) and h = (dl + dr) / 2 carry the common
    // sign, so the limited slope is whichever of them is smaller in magnitude -- no sign to re-insert
    double h = 0.5 * (dl + dr);
    double t = theta * (fabs(dl) < fabs(dr) ? dl : dr);
    double r = fabs(t) < fabs(h) ? t : h;
    return (__double2hiint(dl) ^ __double2hiint(dr)) >= 0 ? r : 0.0;
}

__device__ __forceinline__ double plm_diff(double yl, double y0, double yr, double theta)
{
    return plm_from_differences(y0 - yl, yr - y0, theta);
}

/** TWICE the un-divided PLM difference (stage_tma keeps 2 g: the factor 1/2 of the central slope folds into the
 *  half step of the face states and into the viscous coefficient).  theta2 = 2 theta. */
__device__ __forceinline__ double plm2_from_differences(double dl, double dr, double theta2)
{
    double s = dl + dr;
    double t = theta2 * (fabs(dl) < fabs(dr) ? dl : dr);
    double r = fabs(t) < fabs(s) ? t : s;
    return (__double2hiint(dl) ^ __double2hiint(dr)) >= 0 ? r : 0.0;
}

/** Sound speed squared and geometry-only viscosity factor at a point. */
struct eos_t { double cs2, cs, nu; };

/**
 * cs2_at_position + nu_at_position (scheme.cpp:160-193).
 * y1, y2 out: 1 / sqrt(dr_k^2 + rs^2) for the two bodies (re-used by gravity).
 */
__device__ __forceinline__ double sound_speed_squared(const model_t& M, const stage_t& S, double x, double y, double& y1, double& y2)
{
    double dx1 = x - S.x1, dy1 = y - S.y1, dx2 = x - S.x2, dy2 = y - S.y2;
    double d1 = fma(dx1, dx1, fma(dy1, dy1, M.softening_radius2));
    double d2 = fma(dx2, dx2, fma(dy2, dy2, M.softening_radius2));
    y1 = fast_rsqrt(d1);
    y2 = fast_rsqrt(d2);

    if (M.axisymmetric_cs2)
    {
        return fast_rsqrt(fma(x, x, y * y)) * M.inv_mach2;      // GM / r / M^2
    }
    return fma(S.m1, y1, S.m2 * y2) * M.inv_mach2;              // -(phi1 + phi2) / M^2
}

/** nu_at_position's cutoff profile / constant-nu branch (scheme.cpp:177-193): rare, kept out of line. */
static __device__ __noinline__ double viscosity_slow_path(double nu, double alpha, double alpha_cutoff_radius, double inv_mach, double cs, double r2)
{
    double r = sqrt(r2);
    double profile = alpha_cutoff_radius > 0.0 ? 0.5 * (1.0 + tanh(3.0 * (r - alpha_cutoff_radius))) : 1.0;
    return nu > 0.0 ? profile * nu : profile * alpha * cs * (r * inv_mach);
}

/**
 * Reach of a sink in a2 = dr^2 / (2 s^2).  sink_rate_field (scheme.cpp:117-126) is rate exp(-a2) everywhere; beyond a2 = 40 the
 * factor is below 4.3e-18, so the sink term -u rate exp(-a2) dt (rate dt < 1 for any stable run) is less than a quarter of an
 * ulp of u -- the updated state has the same bits with or without it -- and the cell adds less than 4.3e-18 of a central
 * cell's share to the accretion totals.  (Round 1 used 100; 40 takes 60 % of the area, and of the tiles that run the sink code.)
 */
constexpr double SINK_REACH_A2 = 40.0;

/** sink_rate_field (scheme.cpp:117-126) where it is not negligible: rare, kept out of line. */
static __device__ __noinline__ double sink_weight(double rate, double a2)
{
    return a2 < SINK_REACH_A2 ? rate * exp(-a2) : 0.0;
}

/**
 * The same from the softened squared distances d_k = |x - x_k|^2 + rs^2 to the two bodies and
 * r2 = |x|^2 (the strip kernel tabulates their x- and y-parts per tile row / column).
 */
template<bool FAST>
__device__ __forceinline__ eos_t eos_from_distances(const model_t& M, const stage_t& S, double d1, double d2, double r2)
{
    // FAST: the caller has checked axisymmetric_cs2 == 0, nu == 0, alpha_cutoff_radius == 0 (the defaults): branch-free
    eos_t e;
    e.cs2 = ! FAST && M.axisymmetric_cs2 ? fast_rsqrt(r2) * M.inv_mach2 : fma(S.m1, fast_rsqrt(d1), S.m2 * fast_rsqrt(d2)) * M.inv_mach2;
    e.cs = fast_sqrt(e.cs2);

    if (! FAST && (M.nu > 0.0 || M.alpha_cutoff_radius > 0.0))
    {
        e.nu = viscosity_slow_path(M.nu, M.alpha, M.alpha_cutoff_radius, M.inv_mach, e.cs, r2);
    }
    else
    {
        e.nu = M.alpha * M.inv_mach * fast_sqrt(e.cs2 * r2);
    }
    return e;
}

__device__ __forceinline__ eos_t eos_at_face(const model_t& M, const stage_t& S, double x, double y)
{
    double y1, y2;
    eos_t e;
    e.cs2 = sound_speed_squared(M, S, x, y, y1, y2);
    e.cs = fast_sqrt(e.cs2);
    double r2 = fma(x, x, y * y);

    if (M.nu > 0.0 || M.alpha_cutoff_radius > 0.0)
    {
        e.nu = viscosity_slow_path(M.nu, M.alpha, M.alpha_cutoff_radius, M.inv_mach, e.cs, r2);
    }
    else
    {
        // alpha * sqrt(cs2) * r / M with a single square root: sqrt(cs2 * r^2)
        e.nu = M.alpha * M.inv_mach * fast_sqrt(e.cs2 * r2);
    }
    return e;
}

/**
 * HLLE + viscous flux through a face with normal along `AXIS`:
 * intercell_flux_u (scheme.cpp:268-293) = iso2d::riemann_hlle (physics_iso2d.hpp:488-506)
 * + viscous_flux (scheme.cpp:220-262).
 *
 *   pl, pr       cell-centre primitives left / right of the face
 *   gl, gr       longitudinal gradients (all three components) of the left / right cell
 *   hlx..hry     transverse gradients of vx, vy in the left / right cell
 *   half_step    multiplies gl, gr to reach the face (0.5 * spacing for physical gradients)
 *   visc_scale   multiplies the gradients inside the viscous stress (1 for physical gradients)
 */
/**
 * The HLLE + viscous face flux from the reconstructed states L, R (riemann_hlle, physics_iso2d.hpp:488-506):
 *   F = (ap Fl - am Fr - ap am (Ul - Ur)) / (ap - am),   ap = max(0, vl + cs, vr + cs), am = min(0, vl - cs, vr - cs)
 * regrouped by side.  With wl = ap / (ap - am), wr = -am / (ap - am) (wl + wr = 1) and the mass fluxes
 *   ml = sigma_l wl (vl - am),   mr = sigma_r wr (vr - ap)          (vl - am >= cs, vr - ap <= -cs: no cancellation)
 * the three components are  ml + mr,  ml u_l + mr u_r  (+ cs2 (sigma_l wl + sigma_r wr) along the normal):
 * 15 fp64 instructions after the reciprocal instead of 28.
 *   mu_coef    mu = mu_coef (sigma_l + sigma_r) multiplies d1, d2 (0.5 nu x 0.5 for the face average x the scale of the gradients passed)
 */
template<int AXIS>
__device__ __forceinline__ void hlle_viscous_core(double cs2, double cs, double mu_coef, const prim_t& L, const prim_t& R,
    double d1, double d2, double F[3])
{
    // d1, d2: sums over the two cells of D1 = dx ux - dy uy and D2 = dx uy + dy ux (in the scale mu_coef expects)
    double vl = AXIS == 0 ? L.vx : L.vy;
    double vr = AXIS == 0 ? R.vx : R.vy;
    const bool left_faster = vl > vr;            // one compare serves both extremes
    double ap = dmax0((left_faster ? vl : vr) + cs);        // max(0, vl + cs, vr + cs)
    double am = dmin0((left_faster ? vr : vl) - cs);        // min(0, vl - cs, vr - cs)
    double inv = fast_rcp(ap - am);
    double sl = L.s * (ap * inv), sr = R.s * (-am * inv);      // sigma_l wl, sigma_r wr
    double ml = sl * (vl - am);
    double mr = sr * (vr - ap);
    double pw = (sl + sr) * cs2;
    double f0 = ml + mr;
    double f1 = AXIS == 0 ? fma(ml, L.vx, fma(mr, R.vx, pw)) : fma(ml, L.vx, mr * R.vx);
    double f2 = AXIS == 0 ? fma(ml, L.vy, mr * R.vy) : fma(ml, L.vy, fma(mr, R.vy, pw));

    double mu = mu_coef * (L.s + R.s);
    if (AXIS == 0)
    {
        // tau_xx = mu (dx ux - dy uy), tau_xy = mu (dx uy + dy ux)
        f1 = fma(-mu, d1, f1);
        f2 = fma(-mu, d2, f2);
    }
    else
    {
        // tau_yx = mu (dx uy + dy ux), tau_yy = -mu (dx ux - dy uy)
        f1 = fma(-mu, d2, f1);
        f2 = fma( mu, d1, f2);
    }
    F[0] = f0; F[1] = f1; F[2] = f2;
}

/**
 * HLLE + viscous flux through a face with normal along `AXIS`:
 * intercell_flux_u (scheme.cpp:268-293) = iso2d::riemann_hlle (physics_iso2d.hpp:488-506)
 * + viscous_flux (scheme.cpp:220-262).
 *
 *   pl, pr       cell-centre primitives left / right of the face
 *   gl, gr       longitudinal gradients (all three components) of the left / right cell
 *   hlx..hry     transverse gradients of vx, vy in the left / right cell
 *   half_step    multiplies gl, gr to reach the face (0.5 * spacing for physical gradients)
 *   visc_scale   multiplies the gradients inside the viscous stress (1 for physical gradients)
 */
/**
 * The HLLE + viscous face flux from the reconstructed states L, R (riemann_hlle, physics_iso2d.hpp:488-506):
 *   F = (ap Fl - am Fr - ap am (Ul - Ur)) / (ap - am),   ap = max(0, vl + cs, vr + cs), am = min(0, vl - cs, vr - cs)
 * regrouped by side.  With wl = ap / (ap - am), wr = -am / (ap - am) (wl + wr = 1) and the mass fluxes
 *   ml = sigma_l wl (vl - am),   mr = sigma_r wr (vr - ap)          (vl - am >= cs, vr - ap <= -cs: no cancellation)
 * the three components are  ml + mr,  ml u_l + mr u_r  (+ cs2 (sigma_l wl + sigma_r wr) along the normal):
 * 15 fp64 instructions after the reciprocal instead of 28.
 *   mu_coef    mu = mu_coef (sigma_l + sigma_r) multiplies d1, d2 (0.5 nu x 0.5 for the face average x the scale of the gradients passed)
 */
template<int AXIS>
__device__ __forceinline__ void hlle_viscous_core(double cs2, double cs, double mu_coef, const prim_t& L, const prim_t& R,
    double long_x, double long_y, double tran_x, double tran_y, double F[3])
{
    double vl = AXIS == 0 ? L.vx : L.vy;
    double vr = AXIS == 0 ? R.vx : R.vy;
    const bool left_faster = vl > vr;            // one compare serves both extremes
    double ap = dmax0((left_faster ? vl : vr) + cs);        // max(0, vl + cs, vr + cs)
    double am = dmin0((left_faster ? vr : vl) - cs);        // min(0, vl - cs, vr - cs)
    double inv = fast_rcp(ap - am);
    double sl = L.s * (ap * inv), sr = R.s * (-am * inv);      // sigma_l wl, sigma_r wr
    double ml = sl * (vl - am);
    double mr = sr * (vr - ap);
    double pw = (sl + sr) * cs2;
    double f0 = ml + mr;
    double f1 = AXIS == 0 ? fma(ml, L.vx, fma(mr, R.vx, pw)) : fma(ml, L.vx, mr * R.vx);
    double f2 = AXIS == 0 ? fma(ml, L.vy, mr * R.vy) : fma(ml, L.vy, fma(mr, R.vy, pw));

    double mu = mu_coef * (L.s + R.s);
    if (AXIS == 0)
    {
        // tau_xx = mu (dx ux - dy uy), tau_xy = mu (dx uy + dy ux)
        f1 = fma(-mu, long_x - tran_y, f1);
        f2 = fma(-mu, long_y + tran_x, f2);
    }
    else
    {
        // tau_yx = mu (dx uy + dy ux), tau_yy = -mu (dx ux - dy uy); dx from the transverse set
        f1 = fma(-mu, tran_y + long_x, f1);
        f2 = fma( mu, tran_x - long_y, f2);
    }
    F[0] = f0; F[1] = f1; F[2] = f2;
}

/**
 * HLLE + viscous flux through a face with normal along `AXIS`:
 * intercell_flux_u (scheme.cpp:268-293) = iso2d::riemann_hlle (physics_iso2d.hpp:488-506)
 * + viscous_flux (scheme.cpp:220-262).
 *
 *   pl, pr       cell-centre primitives left / right of the face
 *   gl, gr       longitudinal gradients (all three components) of the left / right cell
 *   hlx..hry     transverse gradients of vx, vy in the left / right cell
 *   half_step    multiplies gl, gr to reach the face (0.5 * spacing for physical gradients)
 *   visc_scale   multiplies the gradients inside the viscous stress (1 for physical gradients)
 */
template<int AXIS>
__device__ __forceinline__ void face_flux(const eos_t& e, prim_t pl, prim_t pr, prim_t gl, prim_t gr,
    double hlx, double hly, double hrx, double hry, double half_step, double visc_scale, double F[3])
{
    prim_t L = {fma(gl.s, half_step, pl.s), fma(gl.vx, half_step, pl.vx), fma(gl.vy, half_step, pl.vy)};
    prim_t R = {fma(-gr.s, half_step, pr.s), fma(-gr.vx, half_step, pr.vx), fma(-gr.vy, half_step, pr.vy)};
    // mu = 0.5 nu (sigma_l + sigma_r); the stresses use face averages 0.5 (g_l + g_r)
    const double long_x = gl.vx + gr.vx, long_y = gl.vy + gr.vy, tran_x = hlx + hrx, tran_y = hly + hry;
    // AXIS 0: longitudinal = d/dx, transverse = d/dy; AXIS 1 the other way round
    const double d1 = AXIS == 0 ? long_x - tran_y : tran_x - long_y;
    const double d2 = AXIS == 0 ? long_y + tran_x : tran_y + long_x;
    hlle_viscous_core<AXIS>(e.cs2, e.cs, (0.25 * visc_scale) * e.nu, L, R, d1, d2, F);
}

/** Running sums behind source_term_total_t (scheme.cpp:22-35, 390-408), before the dt * dA factor. */
using namespace m3b::sums;

/**
 * source_terms_u (scheme.cpp:345-411): gravity of both bodies, both sinks, buffer
 * damping and the density floor, summed in the reference's order.  Adds this
 * cell's contributions to sums[].  Returns y1, y2 (inverse softened distances)
 * for the fused time-step estimate.
 */
template<bool FAST = false, bool WARP_SINKS = false>
__device__ __forceinline__ void source_terms(const model_t& M, const stage_t& S, double x, double y,
    double s, double px, double py, double u0s, double u0x, double u0y, double br,
    double src[3], double sums[NUM_SUMS], double& y1, double& y2, double* warp_sinks = nullptr)
{
    // WARP_SINKS: called by all 32 lanes of a converged warp; the eight sink sums (rarely touched) are
    // accumulated per warp in shared memory (warp_sinks[8]) instead of in 16 registers per thread
    double dx1 = x - S.x1, dy1 = y - S.y1, dx2 = x - S.x2, dy2 = y - S.y2;
    double r1 = fma(dx1, dx1, dy1 * dy1);
    double r2 = fma(dx2, dx2, dy2 * dy2);
    y1 = fast_rsqrt(r1 + M.softening_radius2);
    y2 = fast_rsqrt(r2 + M.softening_radius2);

    // grav_vdot_field * sigma: -dr / (dr^2 + rs^2)^(3/2) G M sigma   (scheme.cpp:85-95, 377-378)
    double k1 = -(y1 * y1) * y1 * S.m1 * s;
    double k2 = -(y2 * y2) * y2 * S.m2 * s;
    double fx1 = dx1 * k1, fy1 = dy1 * k1, fx2 = dx2 * k2, fy2 = dy2 * k2;

    sums[GRV_FX + 0] += fx1;  sums[GRV_FY + 0] += fy1;  sums[GRV_TQ + 0] += fma(x, fy1, -y * fx1);
    sums[GRV_FX + 1] += fx2;  sums[GRV_FY + 1] += fy2;  sums[GRV_TQ + 1] += fma(x, fy2, -y * fx2);

    double a0 = 0.0, a1 = (fx1 + fx2) * S.dt, a2 = (fy1 + fy2) * S.dt;

    // sink_rate_field (scheme.cpp:117-126): rate exp(-dr^2 / (2 s^2)), nothing representable beyond SINK_REACH_A2
    double e1 = r1 * M.sink_inv_2s2, e2 = r2 * M.sink_inv_2s2;

    const bool near_sink = e1 < SINK_REACH_A2 || e2 < SINK_REACH_A2;

    if (WARP_SINKS ? __any_sync(0xffffffffu, near_sink) : near_sink)
    {
        double w1 = sink_weight(M.sink_rate, e1);
        double w2 = sink_weight(M.sink_rate, e2);
        double lz = fma(x, py, -y * px);
        if (WARP_SINKS)
        {
            const double v[8] = {s * w1, s * w2, px * w1, px * w2, py * w1, py * w2, lz * w1, lz * w2};
#ifdef M3B_HOST_EMULATION
            for (int k = 0; k < 8; ++k) warp_sinks[k] += v[k];
#else
            // recursive halving: 9 exchanges instead of 40; lane 4 j ends with the warp's sum of v[j]
            const int lane = threadIdx.x & 31;
            const double v1 = warp_sum8(v, lane);
            if ((lane & 3) == 0) warp_sinks[lane >> 2] += v1;
#endif
        }
        else
        {
            sums[ACC_MASS + 0] += s * w1;   sums[ACC_MASS + 1] += s * w2;
            sums[ACC_PX + 0]   += px * w1;  sums[ACC_PX + 1]   += px * w2;
            sums[ACC_PY + 0]   += py * w1;  sums[ACC_PY + 1]   += py * w2;
            sums[ACC_LZ + 0]   += lz * w1;  sums[ACC_LZ + 1]   += lz * w2;
        }
        double w = -(w1 + w2) * S.dt;
        a0 = fma(s, w, a0);  a1 = fma(px, w, a1);  a2 = fma(py, w, a2);
    }
    // buffer zone (scheme.cpp:384): (U0 - u) * rate * dt; rate is exactly 0 away from the edge, where
    // these terms add exact zeros (FAST: no branch; the rate is non-zero on most of the default domain)
    if (FAST || br != 0.0)
    {
        double b0 = (u0s - s) * br, b1 = (u0x - px) * br, b2 = (u0y - py) * br;
        sums[BUF_M] += b0;
        sums[BUF_L] += fma(x, b2, -y * b1);
        a0 = fma(b0, S.dt, a0);  a1 = fma(b1, S.dt, a1);  a2 = fma(b2, S.dt, a2);
    }
    // density floor (scheme.cpp:385-388): u * 0.01 where sigma < floor
    if (! FAST && s < M.density_floor)      // FAST: the caller has checked density_floor == 0
    {
        a0 = fma(s, 1e-2, a0);  a1 = fma(px, 1e-2, a1);  a2 = fma(py, 1e-2, a2);
    }
    src[0] = a0; src[1] = a1; src[2] = a2;
}

/** to_angmom_fluxes (scheme.cpp:199-214): F(Sr) = x F(px) + y F(py), F(Lz) = x F(py) - y F(px); none through the domain edge. */
template<int AXIS>
__device__ __forceinline__ void to_angmom_fluxes(const model_t& M, double x, double y, double F[3])
{
    const double fpx = F[1], fpy = F[2];
    double flz = x * fpy - y * fpx;
    if (AXIS == 0 && (x == -M.domain_radius || x == M.domain_radius)) flz = 0.0;
    if (AXIS == 1 && (y == -M.domain_radius || y == M.domain_radius)) flz = 0.0;
    F[1] = x * fpx + y * fpy;
    F[2] = flz;
}

/** recover_primitive(Q, x) / to_conserved_per_area(Q, x) (physics_iso2d.hpp:376-417): linear momentum from (Sr, Lz). */
__device__ __forceinline__ void angmom_to_linear(double x, double y, double sr, double lz, double& px, double& py)
{
    const double r2 = x * x + y * y;
    const double a = (sr * x - lz * y) / r2, b = (sr * y + lz * x) / r2;      // (px, py may alias sr, lz)
    px = a;
    py = b;
}

/** The same with one fast reciprocal instead of two divisions (strip kernel). */
__device__ __forceinline__ void angmom_to_linear_fast(double x, double y, double sr, double lz, double& px, double& py)
{
    const double inv = fast_rcp(fma(x, x, y * y));
    const double a = (sr * x - lz * y) * inv, b = (sr * y + lz * x) * inv;
    px = a;
    py = b;
}

/**
 * source_terms_q (scheme.cpp:417-466) for one cell of the angular-momentum-conserving variable set: gravity and sinks
 * in (Sr, Lz) form, buffer towards the initial conserved_q, and the geometric term 2 (E_kin + p) of the radial
 * momentum with its ramp.  The running sums keep the meaning they have for source_terms_u (momentum accreted is the
 * sink term converted back to linear momentum, :446-447), so the host bookkeeping is shared; no work integral.
 */
template<bool WARP_SINKS = false>
__device__ __forceinline__ void source_terms_q(const model_t& M, const stage_t& S, double x, double y,
    double s, double sr, double lz, double vx, double vy, double q0s, double q0r, double q0l, double br,
    double src[3], double sums[NUM_SUMS], double& y1, double& y2, double* warp_sinks = nullptr)
{
    // WARP_SINKS: as in source_terms -- called by all 32 lanes of a converged warp, sink sums per warp in shared memory
    double dx1 = x - S.x1, dy1 = y - S.y1, dx2 = x - S.x2, dy2 = y - S.y2;
    double r1 = fma(dx1, dx1, dy1 * dy1);
    double r2 = fma(dx2, dx2, dy2 * dy2);
    y1 = fast_rsqrt(r1 + M.softening_radius2);
    y2 = fast_rsqrt(r2 + M.softening_radius2);
    double k1 = -(y1 * y1) * y1 * S.m1 * s;
    double k2 = -(y2 * y2) * y2 * S.m2 * s;
    double fx1 = dx1 * k1, fy1 = dy1 * k1, fx2 = dx2 * k2, fy2 = dy2 * k2;
    double tq1 = x * fy1 - y * fx1, tq2 = x * fy2 - y * fx2;

    sums[GRV_FX + 0] += fx1;  sums[GRV_FY + 0] += fy1;  sums[GRV_TQ + 0] += tq1;
    sums[GRV_FX + 1] += fx2;  sums[GRV_FY + 1] += fy2;  sums[GRV_TQ + 1] += tq2;

    double a0 = 0.0;
    double a1 = (x * fx1 + y * fy1) * S.dt + (x * fx2 + y * fy2) * S.dt;
    double a2 = tq1 * S.dt + tq2 * S.dt;

    double e1 = r1 * M.sink_inv_2s2, e2 = r2 * M.sink_inv_2s2;
    const bool near_sink = e1 < SINK_REACH_A2 || e2 < SINK_REACH_A2;
    if (WARP_SINKS ? __any_sync(0xffffffffu, near_sink) : near_sink)
    {
        double w1 = sink_weight(M.sink_rate, e1);
        double w2 = sink_weight(M.sink_rate, e2);
        double px, py;
        angmom_to_linear(x, y, sr, lz, px, py);
        if (WARP_SINKS)
        {
            const double v[8] = {s * w1, s * w2, px * w1, px * w2, py * w1, py * w2, lz * w1, lz * w2};
#ifdef M3B_HOST_EMULATION
            for (int k = 0; k < 8; ++k) warp_sinks[k] += v[k];
#else
            // recursive halving: 9 exchanges instead of 40; lane 4 j ends with the warp's sum of v[j]
            const int lane = threadIdx.x & 31;
            const double v1 = warp_sum8(v, lane);
            if ((lane & 3) == 0) warp_sinks[lane >> 2] += v1;
#endif
        }
        else
        {
            sums[ACC_MASS + 0] += s * w1;   sums[ACC_MASS + 1] += s * w2;
            sums[ACC_PX + 0]   += px * w1;  sums[ACC_PX + 1]   += px * w2;
            sums[ACC_PY + 0]   += py * w1;  sums[ACC_PY + 1]   += py * w2;
            sums[ACC_LZ + 0]   += lz * w1;  sums[ACC_LZ + 1]   += lz * w2;
        }
        double w = -(w1 + w2) * S.dt;
        a0 = fma(s, w, a0);  a1 = fma(sr, w, a1);  a2 = fma(lz, w, a2);
    }
    if (br != 0.0)
    {
        double b0 = (q0s - s) * br, b1 = (q0r - sr) * br, b2 = (q0l - lz) * br;
        sums[BUF_M] += b0;
        sums[BUF_L] += b2;
        a0 = fma(b0, S.dt, a0);  a1 = fma(b1, S.dt, a1);  a2 = fma(b2, S.dt, a2);
    }
    // source_terms_conserved_angmom (physics_iso2d.hpp:277-285): (0, 2 (E_kin + p), 0), ramped up away from the origin
    {
        double cs2 = M.axisymmetric_cs2 ? fast_rsqrt(fma(x, x, y * y)) * M.inv_mach2 : fma(S.m1, y1, S.m2 * y2) * M.inv_mach2;
        // (beyond a = 40 the exponential is below half an ulp of 1: the ramp is exactly 1)
        const double a = (x * x + y * y) / M.gst_suppr_radius2;
        double ramp = a > 40.0 ? 1.0 : 1.0 - exp(-a);
        double ek = 0.5 * s * (vx * vx + vy * vy);
        a1 += (ek + s * cs2) * 2.0 * ramp * S.dt;
    }
    src[0] = a0; src[1] = a1; src[2] = a2;
}

/** max_wavespeed (physics_iso2d.hpp:330-337) with cs2 at the cell centre from y1, y2. */
template<bool FAST = false>
__device__ __forceinline__ double max_wavespeed(const model_t& M, const stage_t& S, double x, double y,
    double y1, double y2, double s, double px, double py)
{
    double cs2 = ! FAST && M.axisymmetric_cs2 ? fast_rsqrt(fma(x, x, y * y)) * M.inv_mach2 : fma(S.m1, y1, S.m2 * y2) * M.inv_mach2;
    double cs = fast_sqrt(cs2);
    // max(|px / s|, |py / s|) = max(|px|, |py|) / s for s > 0 (a product with a positive factor is monotonic: the same bits)
    return fma(dmax(fabs(px), fabs(py)), fast_rcp(s), cs);       // max(|v - cs|, |v + cs|) = |v| + cs for cs >= 0
}


// ---------------------------------------------------------------------------
// stage_tma: the same physics with everything that is constant over a tile, a thread or a stage hoisted
// ---------------------------------------------------------------------------

/** Per-stage constants of stage_tma, derived once per CTA from model_t / stage_t. */
struct strip_consts_t
{
    double m1s, m2s;        // G M_k / Mach^2: cs2 = m1s / sqrt(d1) + m2s / sqrt(d2)   (scheme.cpp:160-175)
    double nm1, nm2;        // -G M_k (gravity, scheme.cpp:85-95)
    double theta2;          // 2 plm_theta
    double dt;
};

/** cs2, cs and the viscous coefficient of a face from tabulated squared distances (FAST: default equation of state).
 *  q2 = (c r)^2 with c = 0.25 * gradient scale * alpha / Mach folded into the table, so that
 *  mu = sqrt(cs2 q2) (sigma_l + sigma_r) = 0.5 nu (sigma_l + sigma_r) x the gradient scale (scheme.cpp:177-193, 284). */
struct eos_face_t { double cs2, cs, mu_coef; };

__device__ __forceinline__ eos_face_t eos_face_fast(const strip_consts_t& C, double d1, double d2, double q2)
{
    eos_face_t e;
    e.cs2 = fma(C.m1s, fast_rsqrt(d1), C.m2s * fast_rsqrt(d2));
    e.cs = fast_sqrt(e.cs2);
    e.mu_coef = fast_sqrt(e.cs2 * q2);
    return e;
}


/** Running sums of one thread of stage_tma (lane <-> column: y and the y-distances to the bodies are constant, so the six
 *  gravity totals of source_term_total_t follow from  S0_k = sum k_k  and  Sx_k = sum x k_k,  k_k = -G M_k sigma / d_k^(3/2):
 *    force_x = Sx_k - x_k S0_k,  force_y = (y - y_k) S0_k,  torque = x_k y S0_k - y_k Sx_k        (scheme.cpp:390-408)
 *  and the ejected angular momentum from  sum x b_py  and  sum b_px). */
struct strip_sums_t
{
    double S0[2], Sx[2];
    double buf_m, buf_xpy, buf_px;
};

/**
 * source_terms_u (scheme.cpp:345-411) for stage_tma: returns acc = u + s (the caller subtracts the flux difference), the
 * inverse softened distances y1, y2 for the fused time-step estimate, and adds to the thread's running sums.
 *   x, dx1, dx2      cell-centre x and its distances to the bodies;  dy1, dy2, y: the thread's constants
 *   d1, d2           softened squared distances (tabulated)
 *   near_sink        warp-uniform: some cell of the tile lies within the sinks' reach (a2 < 100)
 *   has_buffer       warp-uniform: the buffer-zone rate is non-zero somewhere in the tile
 */
template<bool FAST>
__device__ __forceinline__ void source_terms_strip(const model_t& M, const strip_consts_t& C, double x, double y,
    double dx1, double dy1, double dx2, double dy2, double d1, double d2, bool near_sink, bool has_buffer,
    double s, double px, double py, double u0s, double u0x, double u0y, double br,
    double acc[3], strip_sums_t& sums, double& y1, double& y2, double* warp_sinks)
{
    y1 = fast_rsqrt(d1);
    y2 = fast_rsqrt(d2);
    // grav_vdot_field * sigma (scheme.cpp:85-95, 377-378)
    double k1 = ((y1 * y1) * y1) * (s * C.nm1);
    double k2 = ((y2 * y2) * y2) * (s * C.nm2);
    sums.S0[0] += k1;  sums.Sx[0] = fma(x, k1, sums.Sx[0]);
    sums.S0[1] += k2;  sums.Sx[1] = fma(x, k2, sums.Sx[1]);
    double fx = fma(dx1, k1, dx2 * k2), fy = fma(dy1, k1, dy2 * k2);
    double a0 = s, a1 = fma(fx, C.dt, px), a2 = fma(fy, C.dt, py);

    if (near_sink)
    {
        // sink_rate_field (scheme.cpp:117-126): rate exp(-dr^2 / (2 s^2)), nothing representable beyond SINK_REACH_A2.
        // The eight accreted quantities of the warp's 32 cells are folded by recursive halving (9 exchanges instead of 40) and
        // added to the warp's sums in shared memory by the eight lanes that end up with them: no register lives across rows.
        double e1 = fma(dx1, dx1, dy1 * dy1) * M.sink_inv_2s2, e2 = fma(dx2, dx2, dy2 * dy2) * M.sink_inv_2s2;
        if (__any_sync(0xffffffffu, e1 < SINK_REACH_A2 || e2 < SINK_REACH_A2))
        {
            double w1 = sink_weight(M.sink_rate, e1);
            double w2 = sink_weight(M.sink_rate, e2);
            double lz = fma(x, py, -y * px);
            const double v[8] = {s * w1, s * w2, px * w1, px * w2, py * w1, py * w2, lz * w1, lz * w2};
#ifdef M3B_HOST_EMULATION
            for (int k = 0; k < 8; ++k) warp_sinks[k] += v[k];
#else
            const int lane = threadIdx.x & 31;
            const double v1 = warp_sum8(v, lane);
            if ((lane & 3) == 0) warp_sinks[lane >> 2] += v1;
#endif
            double w = -(w1 + w2) * C.dt;
            a0 = fma(s, w, a0);  a1 = fma(px, w, a1);  a2 = fma(py, w, a2);
        }
    }
    if (has_buffer)
    {
        // buffer zone (scheme.cpp:384): (U0 - u) * rate * dt
        double b0 = (u0s - s) * br, b1 = (u0x - px) * br, b2 = (u0y - py) * br;
        sums.buf_m += b0;
        sums.buf_xpy = fma(x, b2, sums.buf_xpy);
        sums.buf_px += b1;
        a0 = fma(b0, C.dt, a0);  a1 = fma(b1, C.dt, a1);  a2 = fma(b2, C.dt, a2);
    }
    if (! FAST && s < M.density_floor)      // density floor (scheme.cpp:385-388): u * 0.01 where sigma < floor
    {
        a0 = fma(s, 1e-2, a0);  a1 = fma(px, 1e-2, a1);  a2 = fma(py, 1e-2, a2);
    }
    acc[0] = a0; acc[1] = a1; acc[2] = a2;
}


/** max_wavespeed with cs2 from the pre-scaled masses (FAST) */
__device__ __forceinline__ double max_wavespeed_fast(const strip_consts_t& C, double y1, double y2, double s, double px, double py)
{
    double cs2 = fma(C.m1s, y1, C.m2s * y2);
    double cs = fast_sqrt(cs2);
    return fma(dmax(fabs(px), fabs(py)), fast_rcp(s), cs);
}

}} // namespace m3b::dev
