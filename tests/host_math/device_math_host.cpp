// tests/host_math/device_math_host.cpp -- TEST INFRASTRUCTURE.  The point-wise device functions of
// mara3_b200/csrc/iso2d_device.cuh compiled by the host compiler (M3B_HOST_EMULATION) and exported with a C ABI, so that
// tests/test_device_math_host.py can check their algebra against the oracle (oracle/libm3b_oracle.so) without a GPU.
#define M3B_HOST_EMULATION 1
#include "../../mara3_b200/csrc/iso2d_device.cuh"

using namespace m3b::dev;

extern "C" {

// riemann_hlle (physics_iso2d.hpp:488-506) through hlle_viscous_core with no viscosity; p = (sigma, vx, vy)
void dm_hlle(const double* pl, const double* pr, double cs2, int axis, double* F)
{
    prim_t L = {pl[0], pl[1], pl[2]}, R = {pr[0], pr[1], pr[2]};
    if (axis == 0) hlle_viscous_core<0>(cs2, std::sqrt(cs2), 0.0, L, R, 0, 0, F);
    else           hlle_viscous_core<1>(cs2, std::sqrt(cs2), 0.0, L, R, 0, 0, F);
}

// intercell flux with viscosity through face_flux (physical gradients: half_step = 0.5 h, visc_scale = 1)
void dm_face_flux(const double* pl, const double* pr, const double* gl, const double* gr, const double* hl, const double* hr,
                  double cs2, double nu, double h, int axis, double* F)
{
    eos_t e = {cs2, std::sqrt(cs2), nu};
    prim_t PL = {pl[0], pl[1], pl[2]}, PR = {pr[0], pr[1], pr[2]}, GL = {gl[0], gl[1], gl[2]}, GR = {gr[0], gr[1], gr[2]};
    if (axis == 0) face_flux<0>(e, PL, PR, GL, GR, hl[0], hl[1], hr[0], hr[1], 0.5 * h, 1.0, F);
    else           face_flux<1>(e, PL, PR, GL, GR, hl[0], hl[1], hr[0], hr[1], 0.5 * h, 1.0, F);
}

double dm_plm(double yl, double y0, double yr, double theta) { return plm_diff(yl, y0, yr, theta); }
double dm_plm2(double yl, double y0, double yr, double theta) { return plm2_from_differences(y0 - yl, yr - y0, 2.0 * theta); }
double dm_max0(double x) { return dmax0(x); }
double dm_min0(double x) { return dmin0(x); }

// source_terms (reference form) and source_terms_strip (regrouped) for one cell; returns src[3] and the 8 + 8 totals
void dm_source_terms(const double* model9, const double* stage, double x, double y, const double* u, const double* u0, double br,
                     double* src, double* sums16)
{
    model_t M = {};
    M.softening_radius2 = model9[0]; M.sink_rate = model9[1]; M.sink_inv_2s2 = model9[2]; M.inv_mach2 = model9[3]; M.inv_mach = model9[4];
    M.alpha = model9[5]; M.density_floor = model9[6];
    stage_t S = {};
    S.dt = stage[0]; S.x1 = stage[1]; S.y1 = stage[2]; S.m1 = stage[3]; S.x2 = stage[4]; S.y2 = stage[5]; S.m2 = stage[6];
    double y1, y2;
    for (int k = 0; k < 16; ++k) sums16[k] = 0.0;
    source_terms<false, false>(M, S, x, y, u[0], u[1], u[2], u0[0], u0[1], u0[2], br, src, sums16, y1, y2);
}

void dm_source_terms_strip(const double* model9, const double* stage, double x, double y, const double* u, const double* u0, double br,
                           double* src, double* sums16)
{
    model_t M = {};
    M.softening_radius2 = model9[0]; M.sink_rate = model9[1]; M.sink_inv_2s2 = model9[2]; M.inv_mach2 = model9[3]; M.inv_mach = model9[4];
    M.alpha = model9[5]; M.density_floor = model9[6];
    stage_t S = {};
    S.dt = stage[0]; S.x1 = stage[1]; S.y1 = stage[2]; S.m1 = stage[3]; S.x2 = stage[4]; S.y2 = stage[5]; S.m2 = stage[6];
    strip_consts_t C = {S.m1 * M.inv_mach2, S.m2 * M.inv_mach2, -S.m1, -S.m2, 0.0, S.dt};
    strip_sums_t sums = {{0.0, 0.0}, {0.0, 0.0}, 0.0, 0.0, 0.0};
    double sinks[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const double dx1 = x - S.x1, dy1 = y - S.y1, dx2 = x - S.x2, dy2 = y - S.y2;
    const double d1 = dx1 * dx1 + (dy1 * dy1 + M.softening_radius2), d2 = dx2 * dx2 + (dy2 * dy2 + M.softening_radius2);
    double acc[3], y1, y2;
    source_terms_strip<false>(M, C, x, y, dx1, dy1, dx2, dy2, d1, d2, true, true, u[0], u[1], u[2], u0[0], u0[1], u0[2], br, acc, sums, y1, y2, sinks);
    for (int q = 0; q < 3; ++q) src[q] = acc[q] - u[q];
    for (int k = 0; k < 8; ++k) sums16[k] = sinks[k];
    // the tile-end reconstruction of stage_tma
    sums16[GRV_FX + 0] = fma(-S.x1, sums.S0[0], sums.Sx[0]);  sums16[GRV_FX + 1] = fma(-S.x2, sums.S0[1], sums.Sx[1]);
    sums16[GRV_FY + 0] = dy1 * sums.S0[0];                    sums16[GRV_FY + 1] = dy2 * sums.S0[1];
    sums16[GRV_TQ + 0] = fma(S.x1 * y, sums.S0[0], -S.y1 * sums.Sx[0]);
    sums16[GRV_TQ + 1] = fma(S.x2 * y, sums.S0[1], -S.y2 * sums.Sx[1]);
    sums16[BUF_M] = sums.buf_m;
    sums16[BUF_L] = fma(-y, sums.buf_px, sums.buf_xpy);
}

// eos_face_fast against eos_from_distances<false>: returns cs2, cs, mu_coef / cvis
void dm_eos(double m1, double m2, double inv_mach, double alpha, double d1, double d2, double r2, double cvis, double* fast3, double* ref3)
{
    model_t M = {};
    M.inv_mach = inv_mach; M.inv_mach2 = inv_mach * inv_mach; M.alpha = alpha;
    stage_t S = {}; S.m1 = m1; S.m2 = m2;
    strip_consts_t C = {m1 * M.inv_mach2, m2 * M.inv_mach2, -m1, -m2, 0.0, 0.0};
    const double cf = cvis * alpha * inv_mach;
    eos_face_t e = eos_face_fast(C, d1, d2, r2 * cf * cf);
    fast3[0] = e.cs2; fast3[1] = e.cs; fast3[2] = e.mu_coef;
    eos_t g = eos_from_distances<false>(M, S, d1, d2, r2);
    ref3[0] = g.cs2; ref3[1] = g.cs; ref3[2] = cvis * g.nu;
}

}
