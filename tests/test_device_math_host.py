"""The point-wise device functions of mara3_b200/csrc/iso2d_device.cuh, compiled by the HOST compiler
(tests/host_math/device_math_host.cpp, M3B_HOST_EMULATION: hardware reciprocal seeds become library calls,
nothing else changes), against the oracle's restatements of the reference routines.  This pins the
regrouped arithmetic of the stage kernels (HLLE by side, doubled PLM differences, gravity totals from two
running sums, pre-scaled equation of state) without a GPU; the GPU parity tests then check the kernels."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
from oracle_util import ORACLE_SO, ROOT

D = C.c_double
P = C.POINTER(C.c_double)


def ptr(a):
    return a.ctypes.data_as(P)


@pytest.fixture(scope="module")
def dm(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host_math") / "libdm.so")
    # -ffp-contract=off: the explicit fma() calls are the device code's FMAs; nothing else is contracted
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", out,
                           os.path.join(ROOT, "tests", "host_math", "device_math_host.cpp")])
    lib = C.CDLL(out)
    lib.dm_hlle.argtypes = [P, P, D, C.c_int, P]
    lib.dm_face_flux.argtypes = [P, P, P, P, P, P, D, D, D, C.c_int, P]
    for f in (lib.dm_plm, lib.dm_plm2):
        f.argtypes = [D, D, D, D]; f.restype = D
    for f in (lib.dm_max0, lib.dm_min0):
        f.argtypes = [D]; f.restype = D
    for f in (lib.dm_source_terms, lib.dm_source_terms_strip):
        f.argtypes = [P, P, D, D, P, P, D, P, P]
    lib.dm_eos.argtypes = [D, D, D, D, D, D, D, D, P, P]
    return lib


@pytest.fixture(scope="module")
def oracle():
    lib = C.CDLL(ORACLE_SO)
    lib.m3o_riemann_hlle.argtypes = [P, P, D, C.c_int, P]
    lib.m3o_plm_gradient.argtypes = [D, D, D, D]; lib.m3o_plm_gradient.restype = D
    return lib


def random_states(rng, n):
    """Left / right primitives and sound speeds covering subsonic faces and faces supersonic in either direction."""
    for k in range(n):
        mach = 10.0 ** rng.uniform(-2, 1.7)
        cs2 = 10.0 ** rng.uniform(-3, 1)
        v = mach * np.sqrt(cs2) * rng.choice([-1.0, 1.0])
        pl = np.array([10.0 ** rng.uniform(-9, -4), v * rng.uniform(0.5, 1.5), rng.normal() * abs(v)])
        pr = np.array([pl[0] * 10.0 ** rng.uniform(-1, 1), v * rng.uniform(0.5, 1.5) + rng.normal() * np.sqrt(cs2), rng.normal() * abs(v)])
        if k % 7 == 0:
            pr = pl.copy()          # identical states: the flux must be the physical flux
        yield pl, pr, cs2


def test_hlle_by_side_equals_riemann_hlle(dm, oracle):
    # physics_iso2d.hpp:488-506 (F = (ap Fl - am Fr - ap am (Ul - Ur)) / (ap - am)) against hlle_viscous_core
    rng = np.random.default_rng(1)
    worst = 0.0
    for pl, pr, cs2 in random_states(rng, 4000):
        for axis in (0, 1):
            got, want = np.zeros(3), np.zeros(3)
            dm.dm_hlle(ptr(pl), ptr(pr), cs2, axis, ptr(got))
            oracle.m3o_riemann_hlle(ptr(pl), ptr(pr), cs2, axis, ptr(want))
            # each component against the size of the terms it is made of (sigma |v| (|v| + cs))
            s = max(pl[0], pr[0]); v = max(np.abs(pl[1:]).max(), np.abs(pr[1:]).max()) + np.sqrt(cs2)
            scale = np.array([s * v, s * v * v, s * v * v])
            worst = max(worst, float((np.abs(got - want) / scale).max()))
    assert worst <= 2e-15, worst


def reference_face_flux(oracle, pl, pr, gl, gr, hl, hr, cs2, nu, h, axis):
    """intercell_flux_u + viscous_flux, scheme.cpp:220-293, with physical gradients."""
    pl_hat = pl + gl * 0.5 * h
    pr_hat = pr - gr * 0.5 * h
    mu = 0.5 * nu * (pl_hat[0] + pr_hat[0])
    F = np.zeros(3)
    oracle.m3o_riemann_hlle(ptr(pl_hat), ptr(pr_hat), cs2, axis, ptr(F))
    if axis == 0:
        dx_ux, dx_uy = 0.5 * (gl[1] + gr[1]), 0.5 * (gl[2] + gr[2])
        dy_ux, dy_uy = 0.5 * (hl[0] + hr[0]), 0.5 * (hl[1] + hr[1])
        return F + np.array([0.0, -mu * (dx_ux - dy_uy), -mu * (dx_uy + dy_ux)])
    dx_ux, dx_uy = 0.5 * (hl[0] + hr[0]), 0.5 * (hl[1] + hr[1])
    dy_ux, dy_uy = 0.5 * (gl[1] + gr[1]), 0.5 * (gl[2] + gr[2])
    return F + np.array([0.0, -mu * (dx_uy + dy_ux), mu * (dx_ux - dy_uy)])


def test_face_flux_with_viscosity(dm, oracle):
    rng = np.random.default_rng(2)
    worst = 0.0
    for pl, pr, cs2 in random_states(rng, 2000):
        h = 10.0 ** rng.uniform(-3, -1)
        v = np.abs(pl[1:]).max() + np.sqrt(cs2)
        gl = np.array([pl[0] * rng.normal() * 0.1, rng.normal() * v * 0.1, rng.normal() * v * 0.1]) / h
        gr = np.array([pr[0] * rng.normal() * 0.1, rng.normal() * v * 0.1, rng.normal() * v * 0.1]) / h
        hl, hr = rng.normal(size=2) * v * 0.1 / h, rng.normal(size=2) * v * 0.1 / h
        nu = 10.0 ** rng.uniform(-4, -1)
        for axis in (0, 1):
            got = np.zeros(3)
            dm.dm_face_flux(ptr(pl), ptr(pr), ptr(gl), ptr(gr), ptr(hl), ptr(hr), cs2, nu, h, axis, ptr(got))
            want = reference_face_flux(oracle, pl, pr, gl, gr, hl, hr, cs2, nu, h, axis)
            s = 1.2 * max(pl[0], pr[0]); vv = 1.2 * v
            scale = np.array([s * vv, s * vv * vv + s * nu * vv * 0.4 / h, s * vv * vv + s * nu * vv * 0.4 / h])
            worst = max(worst, float((np.abs(got - want) / scale).max()))
    assert worst <= 4e-15, worst


def test_plm_differences(dm, oracle):
    # math_interpolation.hpp:85-94; the device keeps the un-divided difference (stage_strip) or twice it (stage_tma)
    rng = np.random.default_rng(3)
    cases = [(0.0, 0.0, 0.0), (1.0, 1.0, 2.0), (1.0, 2.0, 2.0), (1.0, 2.0, 1.0), (-1.0, 0.0, 1.0), (0.0, -0.0, 0.0), (3.0, 1.0, -2.0)]
    cases += [tuple(rng.normal(size=3)) for _ in range(5000)]
    cases += [(a, a + 1e-17, a + 2e-17) for a in rng.normal(size=50)]
    for theta in (0.0, 1.0, 1.8, 2.0):
        for yl, y0, yr in cases:
            want = oracle.m3o_plm_gradient(yl, y0, yr, theta)
            tol = 4e-16 * (abs(y0 - yl) + abs(yr - y0))       # b = (yr - yl) / 2 against (dl + dr) / 2: one rounding of the sum
            assert abs(dm.dm_plm(yl, y0, yr, theta) - want) <= tol
            assert abs(dm.dm_plm2(yl, y0, yr, theta) - 2.0 * want) <= 2 * tol


def test_integer_pipe_max0_min0(dm):
    for x in (0.0, -0.0, 1.5, -1.5, 1e-300, -1e-300, 1e300, -1e300, 4.9e-324, -4.9e-324):
        assert dm.dm_max0(x) == max(0.0, x) and dm.dm_min0(x) == min(0.0, x)
        assert not np.signbit(dm.dm_max0(x))


def test_source_terms_strip_equals_source_terms(dm):
    """The regrouped source terms of stage_tma (u + s folded, gravity totals from S0 = sum k and Sx = sum x k) against
    the reference-order source_terms of the other kernels, cell by cell (scheme.cpp:345-411)."""
    rng = np.random.default_rng(4)
    model = np.array([0.05 ** 2, 1.0, 1.0 / 0.05 ** 2 / 2.0, 1e-2, 1e-1, 0.1, 0.0, 0.0, 0.0])
    worst_src = worst_sum = 0.0
    for k in range(3000):
        stage = np.array([10.0 ** rng.uniform(-4, -2), 0.5 * np.cos(k), 0.5 * np.sin(k), 0.5, -0.5 * np.cos(k), -0.5 * np.sin(k), 0.5])
        r = 10.0 ** rng.uniform(-2, 1)
        phi = rng.uniform(0, 2 * np.pi)
        x, y = (r * np.cos(phi), r * np.sin(phi)) if k % 3 else (stage[1] + rng.normal() * 0.05, stage[2] + rng.normal() * 0.05)
        s = 10.0 ** rng.uniform(-9, -5)
        u = np.array([s, s * rng.normal() * 3, s * rng.normal() * 3])
        u0 = u * rng.uniform(0.5, 1.5, size=3)
        br = rng.choice([0.0, 10.0 ** rng.uniform(-3, 1)])
        a, sa, b, sb = np.zeros(3), np.zeros(16), np.zeros(3), np.zeros(16)
        dm.dm_source_terms(ptr(model), ptr(stage), x, y, ptr(u), ptr(u0), br, ptr(a), ptr(sa))
        dm.dm_source_terms_strip(ptr(model), ptr(stage), x, y, ptr(u), ptr(u0), br, ptr(b), ptr(sb))
        # the strip form returns (u + s) - u: its rounding is relative to max(|u|, |s|), which is what the update adds s to anyway
        worst_src = max(worst_src, float((np.abs(a - b) / np.maximum(np.abs(u), np.abs(a))).max()))
        # totals: against the size of the terms they are differences of (|x| + |x_k|) |f|
        fscale = max(np.abs(sa[8:12]).max(), 1e-300)
        scale = np.array([np.abs(sa[k2]).max() + 1e-300 for k2 in range(8)] + [fscale] * 4 + [(abs(x) + abs(y) + 1.0) * fscale] * 2
                         + [abs(sa[14]) + 1e-300, (abs(x) + abs(y)) * np.abs(u).max() * max(br, 1e-300) + 1e-300])
        worst_sum = max(worst_sum, float((np.abs(sa - sb) / scale).max()))
    assert worst_src <= 1e-15, worst_src
    assert worst_sum <= 1e-13, worst_sum


def test_sink_reach_leaves_the_state_untouched():
    """iso2d_device.cuh: SINK_REACH_A2 = 40.  The reference applies rate * exp(-a2) everywhere (scheme.cpp:117-126); beyond
    a2 = 40 the sink term -u * rate * exp(-a2) * dt is less than a quarter of an ulp of u for every stable step (rate * dt < 1),
    so u + term == u in floating point, and a cell there adds less than 4.3e-18 of a central cell's share to the totals."""
    header = open(os.path.join(ROOT, "mara3_b200", "csrc", "iso2d_device.cuh")).read()
    assert "constexpr double SINK_REACH_A2 = 40.0;" in header
    factor = np.exp(-40.0)
    assert factor < 4.3e-18 and factor < 0.25 * np.finfo(np.float64).eps
    rng = np.random.default_rng(11)
    for u in 10.0 ** rng.uniform(-12, 3, size=200):
        for rate_dt in (1.0, 0.5, 1e-3):
            assert u + (-u * rate_dt * factor) == u


def test_eos_face_fast(dm):
    # cs2_at_position / nu_at_position (scheme.cpp:160-193) from pre-scaled masses and the pre-scaled r^2 table
    rng = np.random.default_rng(5)
    worst = 0.0
    for _ in range(3000):
        d1, d2, r2 = 10.0 ** rng.uniform(-2.6, 2.5), 10.0 ** rng.uniform(-2.6, 2.5), 10.0 ** rng.uniform(-4, 2.5)
        m1 = rng.uniform(0.1, 0.9)
        fast, ref = np.zeros(3), np.zeros(3)
        dm.dm_eos(m1, 1.0 - m1, 1.0 / rng.uniform(5, 40), 0.1, d1, d2, r2, 0.125 * 10.0 ** rng.uniform(1, 3), ptr(fast), ptr(ref))
        worst = max(worst, float((np.abs(fast - ref) / np.abs(ref)).max()))
    assert worst <= 2e-15, worst
