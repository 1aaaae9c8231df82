"""GPU (-m gpu): the CUDA path, called through the C ABI, against the plain-C oracle and the
reference's golden vectors.

Tolerances (BASELINE.json north_star): fp64; per cell <= 1e-12 relative after one step, measured
as |gpu - ref| / max|ref| over the cell's (block, field) because momenta cross zero; diagnostics
(the time-series scalars) <= 1e-9 after 100 steps.  Tree layout / indexing is bit-exact and is
covered on the CPU in test_host_layer.py."""
import numpy as np
import pytest
import mara3_b200 as m3
from conftest import load_golden, block_rel_err, block_rel_err_q
from oracle_util import OracleMesh, OracleSolution, SCALAR_NAMES

pytestmark = pytest.mark.gpu

CELL_TOL = 1e-12
DIAG_TOL = 1e-9


def scalar_errors(got, want, elapsed):
    """Error of each of the 43 scalars relative to its natural scale.  Accumulators that vanish by
    symmetry, orbital-element differences (O(1) elements subtracted) and the work integral (difference
    of O(0.1) kinetic energies) carry absolute rounding noise, so each gets an absolute floor."""
    disk = 1e-3 * max(elapsed, 1e-3)            # disk_mass (config default) x elapsed time, GM = a = 1
    floors = np.full(43, 0.0)
    floors[3:13] = disk                         # mass / angular momentum / torque / ejected accumulators
    floors[9:11] = 1e-4                         # work: difference of two O(0.1) kinetic energies
    floors[13:43] = 1.0                         # orbital elements and their accumulated differences: O(1) quantities
    if want[42] == 0.0:
        # circular binary: argument of periapse, periapse time and the eccentricity itself (sqrt of a
        # rounding-level number) of the PERTURBED orbits are ill-conditioned -- in the reference too
        floors[[13, 14, 22, 23, 24, 32]] = np.inf
    floors[0] = elapsed
    floors[1:3] = 1.0
    return np.abs(got - want) / np.maximum(np.abs(want), floors)


def assert_scalars(got, want, tol, elapsed):
    err = scalar_errors(got, want, elapsed)
    k = int(err.argmax())
    assert err[k] <= tol, (SCALAR_NAMES[k], got[k], want[k], err[k])


def make_pair(cfg, **kw):
    solver = m3.Solver(cfg, **kw)
    return solver, solver.create_solution(), OracleSolution(OracleMesh(cfg))


# ---------------------------------------------------------------------------------------------
# one stage / one step against the oracle
# ---------------------------------------------------------------------------------------------
STAGE_CASES = [
    (dict(depth=2, block_size=64, domain_radius=6.0), False),      # config 1 grid, warp-strip kernel (16x32 tiles)
    (dict(depth=2, block_size=64, domain_radius=6.0), "tiled"),    # same through the generic tiled kernel
    (dict(depth=2, block_size=64, domain_radius=6.0), True),       # same through the any-tree kernels
    (dict(depth=5, block_size=32), False),                         # nested: strip kernel, jump blocks through its JUMP variant
    (dict(depth=4, block_size=64), False),                         # default nested tree in 64^2 blocks (the C4 block size)
    (dict(depth=4, block_size=32, eccentricity=0.3, mass_ratio=0.5, nu=0.01, alpha_cutoff_radius=1.0,
          density_floor=1e-2), False),                              # JUMP variant without the branch-free equation of state
    (dict(depth=3, block_size=128), False),                        # 128^2 blocks: tiles without any block side, run-time block size in both variants
    (dict(depth=4, block_size=24), False),                         # default nested tree, fused 12x24 + jumps
    (dict(depth=6, block_size=16), False),                         # five levels, fused 16x16 + jumps
    (dict(depth=5, block_size=8, focus_factor=3.0), False),        # fused 8x8 + jumps
    (dict(depth=3, block_size=10), False),                         # no tile divides 10: general path only
    (dict(depth=3, block_size=8, eccentricity=0.3, mass_ratio=0.5, nu=0.01, alpha_cutoff_radius=1.0,
          begin_live_binary=0.0, density_floor=1e-2, mdot=1e-4), False),
    (dict(depth=3, block_size=16, axisymmetric_cs2=1, counter_rotate=1, focus_factor=1e3), False),
    (dict(depth=3, block_size=16, alpha=0.0, sink_rate=5.0, sink_radius=0.1, softening_radius=0.1,
          buffer_damping_rate=0.0, mach_number=5.0, plm_theta=1.0, focus_factor=1e3), False),
    # angular-momentum-conserving variables (advance_q, conserve_linear_p=0): state = conserved_q, any-tree kernels
    (dict(depth=3, block_size=16, conserve_linear_p=0, fixed_dt=1), False),                     # nested, 16x16 tiles
    (dict(depth=3, block_size=8, conserve_linear_p=0, fixed_dt=1), False),                      # nested, one CTA per block
    (dict(depth=2, block_size=32, conserve_linear_p=0, fixed_dt=1, domain_radius=6.0, focus_factor=1e3), False),    # strip kernel, QMODE
    (dict(depth=2, block_size=32, conserve_linear_p=0, fixed_dt=1, domain_radius=6.0, focus_factor=1e3), True),     # same through the any-tree kernels
    (dict(depth=4, block_size=32, conserve_linear_p=0, fixed_dt=1), False),                     # nested: strip QMODE, JUMP + QMODE at the jumps
    (dict(depth=4, block_size=32, conserve_linear_p=0, fixed_dt=1), True),                      # same through the any-tree kernels
    (dict(depth=4, block_size=64, conserve_linear_p=0, fixed_dt=1, axisymmetric_cs2=1), False), # 64^2 blocks, JUMP + QMODE without the fast equation of state
    (dict(depth=2, block_size=64, conserve_linear_p=0, fixed_dt=1, eccentricity=0.2, mass_ratio=0.5, nu=0.01,
          begin_live_binary=0.0, focus_factor=1e3), False),                                      # strip QMODE, 64^2 blocks, general equation of state
    (dict(depth=2, block_size=16, domain_radius=6.0, conserve_linear_p=0, fixed_dt=1, rk_order=1, eccentricity=0.2,
          mass_ratio=0.5, begin_live_binary=0.0, nu=0.01), False),
]


@pytest.mark.parametrize("cfg,general_only", STAGE_CASES)
def test_advance_matches_oracle(cfg, general_only):
    kw = dict(tiled_kernel=True) if general_only == "tiled" else dict(general_only=general_only)
    solver, u, o = make_pair(cfg, **kw)
    dt_o = 0.4 * o.maximum_timestep()
    dt_g = 0.4 * solver.maximum_timestep(u)
    assert abs(dt_g - dt_o) <= 1e-14 * dt_o
    assert np.array_equal(u.conserved_u, o.conserved_u)             # create_solution: bit-exact initial data
    o1, status = o.advance(dt_o)
    assert status == 0
    g1 = solver.advance(u, dt_o)
    # (conserved_q: Sr nearly cancels in a Keplerian disk and is measured against r |p|, see block_rel_err_q)
    rel_err = block_rel_err if cfg.get("conserve_linear_p", 1) else block_rel_err_q
    assert rel_err(g1.conserved_u, o1.conserved_u) <= CELL_TOL
    assert_scalars(g1.scalars, o1.scalars, 1e-11, dt_o)
    assert np.array_equal(u.conserved_u, o.conserved_u)             # the input solution is untouched (value semantics)
    # safe mode: theta = 0 (piecewise constant)
    o2, _ = o1.advance(0.1 * dt_o, safe_mode=True)
    g2 = solver.advance(g1, 0.1 * dt_o, safe_mode=True)
    assert rel_err(g2.conserved_u, o2.conserved_u) <= CELL_TOL


@pytest.mark.parametrize("seed", [1, 2])
@pytest.mark.parametrize("cfg", [dict(depth=4, block_size=16), dict(depth=2, block_size=32, focus_factor=1e3), dict(depth=4, block_size=32)])
def test_advance_on_rough_seeded_states(cfg, seed):
    """Seeded multiplicative noise makes every limiter / wave-speed branch fire."""
    solver, u, o = make_pair(cfg)
    rng = np.random.default_rng(seed)
    U = o.conserved_u
    U[:, 0] *= rng.uniform(0.5, 1.5, U[:, 0].shape)
    U[:, 1:] *= rng.uniform(-1.0, 2.0, U[:, 1:].shape)
    o.conserved_u = U
    u.conserved_u = U
    dt = 0.2 * o.maximum_timestep()
    assert abs(0.2 * solver.maximum_timestep(u) - dt) <= 1e-14 * dt
    o1, status = o.advance(dt)
    try:
        g1 = solver.advance(u, dt)
        assert status == 0
    except m3.NegativeDensity as e:
        assert status == 1
        g1 = e.solution
    assert block_rel_err(g1.conserved_u, o1.conserved_u) <= CELL_TOL


def test_jump_strip_agrees_with_any_tree_kernels(monkeypatch):
    """Blocks at refinement jumps: stage_strip<.., JUMP> against the 16 x 16 any-tree kernels it replaces
    (M3B_JUMP_STRIP=0), two RK2 steps on a nested tree; both are within the oracle tolerance of each other."""
    cfg = dict(depth=5, block_size=64)
    new = m3.Solver(cfg)
    monkeypatch.setenv("M3B_JUMP_STRIP", "0")
    old = m3.Solver(cfg)
    monkeypatch.delenv("M3B_JUMP_STRIP")
    assert new.num_regular_blocks < new.num_blocks          # the tree has jumps
    un, uo = new.create_solution(), old.create_solution()
    for _ in range(2):
        dn, _ = new.next_solution(un)
        do, _ = old.next_solution(uo)
        assert abs(dn - do) <= 1e-13 * do
    assert block_rel_err(un.conserved_u, uo.conserved_u) <= 2 * CELL_TOL
    assert_scalars(un.scalars, uo.scalars, 1e-10, un.time)


def test_jump_block_tiles_through_the_regular_kernel_agree_with_the_jump_kernel(monkeypatch):
    """A tile of a block at a refinement jump that touches same-level leaves only rides with the regular blocks' tiles in the
    persistent kernel (stage_tma) -- the default -- instead of stage_strip<.., JUMP> (M3B_SPLIT_JUMP=0: every tile of such a
    block through JUMP).  Both see the same cells; two RK2 steps of a nested tree, and one stage against the oracle."""
    cfg = dict(depth=5, block_size=64)
    new = m3.Solver(cfg)
    monkeypatch.setenv("M3B_SPLIT_JUMP", "0")
    old = m3.Solver(cfg)
    monkeypatch.delenv("M3B_SPLIT_JUMP")
    assert new.num_regular_blocks < new.num_blocks          # the tree has jumps
    un, uo = new.create_solution(), old.create_solution()
    for _ in range(2):
        dn, _ = new.next_solution(un)
        do, _ = old.next_solution(uo)
        assert abs(dn - do) <= 1e-13 * do
    assert block_rel_err(un.conserved_u, uo.conserved_u) <= 2 * CELL_TOL
    assert_scalars(un.scalars, uo.scalars, 1e-10, un.time)
    # the launch counts differ only if the split is active somewhere: the same kernels run, over different tile lists
    solver, u, o = make_pair(cfg)
    dt = solver.maximum_timestep(u)
    o1, status = o.advance(dt)
    assert status == 0
    assert block_rel_err(solver.advance(u, dt).conserved_u, o1.conserved_u) <= CELL_TOL


@pytest.mark.parametrize("name", ["nested_d3_n8", "uniform_d2_n16", "live_ecc_d3_n8", "rk1_axisym_d3_n8",
                                  "angmom_nested_d3_n8", "angmom_live_rk1_d2_n16"])
def test_steps_match_reference_golden_vectors(name):
    """tests/golden/*.npz were produced by the reference's own compiled code."""
    g = load_golden(name)
    solver = m3.Solver(g["config"])
    u = solver.create_solution()
    for n in range(1, max(g["steps"]) + 1):
        dt, fell_back = solver.next_solution(u)
        assert not fell_back
        assert abs(dt - g["dt_history"][n - 1]) <= 1e-12 * dt
        if n in g["steps"]:
            assert block_rel_err(u.conserved_u, g[f"step{n}_conserved_u"]) <= n * CELL_TOL
            assert_scalars(u.scalars, g[f"step{n}_scalars"], 1e-10, u.time)
            assert u.iteration == (n, 1)


def test_rk2_from_public_operators_equals_next_solution():
    """s0 * 1/2 + advance(advance(s0)) * 1/2 built from m3b_advance / m3b_solution_combine."""
    cfg = dict(depth=4, block_size=16)
    solver, u, o = make_pair(cfg)
    dt = 0.4 * solver.maximum_timestep(u)
    s1 = solver.advance(u, dt)
    s2 = solver.advance(s1, dt)
    manual = solver.combine(u, s2, 0.5)
    dt_used, _ = solver.next_solution(u)
    assert abs(dt_used - dt) <= 1e-15 * dt
    assert block_rel_err(manual.conserved_u, u.conserved_u) <= 1e-14
    assert np.allclose(manual.scalars, u.scalars, rtol=1e-12, atol=1e-18)
    o.next_solution()
    assert block_rel_err(u.conserved_u, o.conserved_u) <= CELL_TOL


def test_host_buffer_entry_points_equal_device_path():
    cfg = dict(depth=3, block_size=16)
    solver, u, o = make_pair(cfg)
    U0, S0 = u.conserved_u, u.scalars
    dt = 0.4 * solver.maximum_timestep(u)
    U1, S1 = solver.advance_host(U0, S0, dt)
    g1 = solver.advance(u, dt)
    assert np.array_equal(U1, g1.conserved_u) and np.array_equal(S1, g1.scalars)
    U2, S2, dt2, fb = solver.next_solution_host(U0, S0)
    solver.next_solution(u)
    assert np.array_equal(U2, u.conserved_u) and np.array_equal(S2, u.scalars) and not fb
    assert abs(dt2 - dt) <= 1e-15 * dt


# ---------------------------------------------------------------------------------------------
# failure path: negative density -> safe-mode retry (subprog_binary.cpp:285-292)
# ---------------------------------------------------------------------------------------------
def test_safe_mode_retry_matches_reference_behaviour():
    """Config 1 with the default domain_radius runs into negative densities at steps 23-33
    (SURVEY.md section 5); the retry must fire on the same steps and give the same state."""
    cfg = dict(depth=2, block_size=64)
    solver, u, o = make_pair(cfg)
    gpu_fb, cpu_fb = [], []
    for n in range(1, 37):
        dt_g, fb_g = solver.next_solution(u)
        dt_o, fb_o = o.next_solution()
        if fb_g:
            gpu_fb.append(n)
            assert any(m.startswith("negative density") for m in solver.messages())
        if fb_o:
            cpu_fb.append(n)
        assert abs(dt_g - dt_o) <= 1e-10 * dt_o
    assert gpu_fb == cpu_fb and len(gpu_fb) >= 5
    assert block_rel_err(u.conserved_u, o.conserved_u) <= 1e-9
    assert_scalars(u.scalars, o.scalars, DIAG_TOL, u.time)


def test_advance_reports_negative_density_like_validate_u():
    cfg = dict(depth=2, block_size=16, focus_factor=1e3)
    solver, u, o = make_pair(cfg)
    dt = 40.0 * solver.maximum_timestep(u)          # grossly unstable step
    o1, status = o.advance(dt)
    assert status == 1
    with pytest.raises(m3.NegativeDensity) as info:
        solver.advance(u, dt)
    assert info.value.lines and info.value.lines[0].startswith("negative density")
    assert block_rel_err(info.value.solution.conserved_u, o1.conserved_u) <= 1e-11


# ---------------------------------------------------------------------------------------------
# 100-step diagnostics
# ---------------------------------------------------------------------------------------------
def test_diagnostics_after_100_steps():
    cfg = dict(depth=2, block_size=64, domain_radius=6.0)      # config 1 grid without retries
    solver, u, o = make_pair(cfg)
    for n in range(100):
        solver.next_solution(u)
        o.next_solution()
    assert abs(u.time - o.time) <= DIAG_TOL * o.time
    assert o.time == pytest.approx(0.5743490877852454, rel=1e-13)         # survey known answer
    assert_scalars(u.scalars, o.scalars, DIAG_TOL, o.time)
    # time-series diagnostics: disk mass and angular momentum (subprog_binary_diagnostics.cpp:19-42)
    dA, xc = solver.cell_areas, solver.cell_centers
    for U_g, U_o in [(u.conserved_u, o.conserved_u)]:
        mass_g, mass_o = (U_g[:, 0] * dA).sum(), (U_o[:, 0] * dA).sum()
        lz_g = ((xc[:, 0] * U_g[:, 2] - xc[:, 1] * U_g[:, 1]) * dA).sum()
        lz_o = ((xc[:, 0] * U_o[:, 2] - xc[:, 1] * U_o[:, 1]) * dA).sum()
        assert abs(mass_g - mass_o) <= DIAG_TOL * abs(mass_o)
        assert abs(lz_g - lz_o) <= DIAG_TOL * abs(lz_o)
    assert U_o[:, 0].sum() == pytest.approx(0.43661329337982169, rel=1e-12)
    assert block_rel_err(u.conserved_u, o.conserved_u) <= 1e-9


def test_nested_tree_100_steps():
    cfg = dict(depth=4, block_size=16)
    solver, u, o = make_pair(cfg)
    for n in range(100):
        _, fb_g = solver.next_solution(u)
        _, fb_o = o.next_solution()
        assert fb_g == fb_o
    assert_scalars(u.scalars, o.scalars, DIAG_TOL, o.time)
    assert block_rel_err(u.conserved_u, o.conserved_u) <= 1e-9


# ---------------------------------------------------------------------------------------------
# BASELINE.json sizes: direct comparison where the oracle finishes in seconds, properties beyond
# ---------------------------------------------------------------------------------------------
def total_mass(solver, U):
    return float((U[:, 0] * solver.cell_areas).sum())


def test_config2_1024sq_stage_against_oracle():
    cfg = dict(depth=4, block_size=64, focus_factor=1e3, mach_number=10.0)
    solver, u, o = make_pair(cfg)
    assert solver.num_cells == 1024 ** 2 and solver.num_regular_blocks == 256
    dt = 0.4 * o.maximum_timestep()
    assert abs(0.4 * solver.maximum_timestep(u) - dt) <= 1e-14 * dt
    o1, _ = o.advance(dt)
    g1 = solver.advance(u, dt)
    assert block_rel_err(g1.conserved_u, o1.conserved_u) <= CELL_TOL
    assert_scalars(g1.scalars, o1.scalars, 1e-11, dt)


def test_config3_4096sq_stage_against_oracle():
    """BASELINE.json configs[2], the workload bench.py times at every N: one RK stage of the whole 4096^2 grid against the
    oracle, cell by cell (the oracle needs ~20 s for it; full steps at this size are covered by the conservation test below)."""
    cfg = dict(depth=6, block_size=64, focus_factor=1e3, mach_number=10.0)
    solver, u, o = make_pair(cfg)
    assert solver.num_cells == 4096 ** 2 and solver.num_regular_blocks == 4096
    dt = 0.4 * o.maximum_timestep()
    assert abs(0.4 * solver.maximum_timestep(u) - dt) <= 1e-14 * dt
    o1, status = o.advance(dt)
    assert status == 0
    g1 = solver.advance(u, dt)
    assert block_rel_err(g1.conserved_u, o1.conserved_u) <= CELL_TOL
    assert_scalars(g1.scalars, o1.scalars, 1e-11, dt)


@pytest.mark.parametrize("extra", [dict(), dict(conserve_linear_p=0, fixed_dt=1)])
def test_config4_depth8_step_against_oracle(extra):
    """BASELINE.json configs[3] as bench.py's `c4` runs it (depth 8, 64^2 blocks: 424 leaves on six levels, 77 % of them at
    refinement jumps): one full step -- dt rule, both RK stages, combination -- against the oracle, in the linear-momentum
    variables (advance_u, scheme.cpp:790-904) and in the angular-momentum variables (advance_q, :906-1020)."""
    cfg = dict(depth=8, block_size=64, **extra)
    solver, u, o = make_pair(cfg)
    assert solver.num_blocks == 424 and 0 < solver.num_regular_blocks < 424
    dt_o, fb_o = o.next_solution()
    dt_g, fb_g = solver.next_solution(u)
    assert fb_o == fb_g and abs(dt_g - dt_o) <= 1e-13 * dt_o
    rel_err = block_rel_err_q if extra else block_rel_err
    assert rel_err(u.conserved_u, o.conserved_u) <= CELL_TOL
    assert_scalars(u.scalars, o.scalars, 1e-10, o.time)


def test_nested_bench_workload_stage_against_oracle():
    """bench.py's `c4x` (depth 8, focus_factor 12: 4552 leaves of 64^2 on levels 4-8, the nested workload with load for
    8 GPUs): one RK stage against the oracle."""
    cfg = dict(depth=8, block_size=64, focus_factor=12.0)
    solver, u, o = make_pair(cfg)
    assert solver.num_blocks == 4552
    dt = 0.4 * o.maximum_timestep()
    assert abs(0.4 * solver.maximum_timestep(u) - dt) <= 1e-14 * dt
    o1, status = o.advance(dt)
    assert status == 0
    g1 = solver.advance(u, dt)
    assert block_rel_err(g1.conserved_u, o1.conserved_u) <= CELL_TOL


def test_persistent_and_one_shot_stage_kernels_agree(monkeypatch):
    """stage_tma (persistent, cp.async staging, regrouped arithmetic) against stage_strip (M3B_STAGE=strip) over three steps of
    the 1024^2 grid: two implementations of the same stage that share neither the tile loader nor the order of the flux
    arithmetic; both lie within the oracle tolerance of each other."""
    cfg = dict(depth=4, block_size=64, focus_factor=1e3, mach_number=10.0)
    new = m3.Solver(cfg)
    monkeypatch.setenv("M3B_STAGE", "strip")
    old = m3.Solver(cfg)
    monkeypatch.delenv("M3B_STAGE")
    ua, ub = new.create_solution(), old.create_solution()
    for _ in range(3):
        dta, _ = new.next_solution(ua)
        dtb, _ = old.next_solution(ub)
        assert abs(dta - dtb) <= 1e-13 * dtb
    assert block_rel_err(ua.conserved_u, ub.conserved_u) <= 3e-12
    assert_scalars(ua.scalars, ub.scalars, 1e-10, ub.time)


@pytest.mark.parametrize("cfg", [
    dict(depth=4, block_size=64, focus_factor=1e3),     # config 2: 1024^2 uniform
    dict(depth=6, block_size=64, focus_factor=1e3),     # config 3: 4096^2 uniform
    dict(depth=6, block_size=64),                       # config 4: nested, level jumps (136 blocks)
])
def test_full_size_conservation_and_symmetry(cfg):
    """Size-independent properties: (1) the flux-difference update is conservative, also across
    refinement jumps (flux correction), so the disk mass changes exactly by what the sinks and the
    buffer removed; (2) an equal-mass circular binary in a symmetric disk keeps the solution
    symmetric under rotation by pi."""
    solver = m3.Solver(cfg)
    u = solver.create_solution()
    m0 = total_mass(solver, u.conserved_u)
    nsteps = 4
    for _ in range(nsteps):
        solver.next_solution(u)
    U = u.conserved_u
    s = u.scalars
    removed = s[3] + s[4] + s[11]                       # mass_accreted_on[0,1] + mass_ejected
    m1 = total_mass(solver, U)
    assert abs((m0 - m1) - removed) <= 2e-13 * m0
    assert u.iteration == (nsteps, 1)
    # rotation by pi maps block (l, i, j) to (l, n-1-i, n-1-j) and flips each block
    idx = solver.tree_index
    lookup = {tuple(r): b for b, r in enumerate(idx)}
    partner = np.array([lookup[(l, (1 << l) - 1 - i, (1 << l) - 1 - j)] for l, i, j in idx])
    R = U[partner][:, :, ::-1, ::-1]
    scale = np.abs(U).max(axis=(2, 3), keepdims=True)
    assert (np.abs(U[:, 0] - R[:, 0]) / scale[:, 0]).max() <= 1e-9
    assert (np.abs(U[:, 1] + R[:, 1]) / scale[:, 1]).max() <= 1e-9
    assert (np.abs(U[:, 2] + R[:, 2]) / scale[:, 2]).max() <= 1e-9
    assert abs(s[3] - s[4]) <= 1e-9 * abs(s[3])         # both bodies accrete equally


def test_fused_and_general_kernels_agree_at_size():
    cfg = dict(depth=4, block_size=64, focus_factor=1e3)
    a = m3.Solver(cfg)
    ua = a.create_solution()
    for _ in range(2):
        a.next_solution(ua)
    for kw in (dict(general_only=True), dict(tiled_kernel=True)):
        b = m3.Solver(cfg, **kw)
        ub = b.create_solution()
        for _ in range(2):
            b.next_solution(ub)
        assert block_rel_err(ua.conserved_u, ub.conserved_u) <= 1e-13
        assert np.allclose(ua.scalars, ub.scalars, rtol=1e-10, atol=1e-16)


def test_pipelined_stepping_equals_step_by_step():
    """next_solution queues the following step ahead of the host; with that switched off every call
    starts and finishes one step.  Same fields to rounding of the device-side sin/cos, same retries."""
    cfg = dict(depth=2, block_size=64)
    a, b = m3.Solver(cfg), m3.Solver(cfg)
    b.set_pipelining(False)
    ua, ub = a.create_solution(), b.create_solution()
    fa, fb = [], []
    for n in range(30):
        dta, ra = a.next_solution(ua)
        dtb, rb = b.next_solution(ub)
        fa.append(ra)
        fb.append(rb)
        assert abs(dta - dtb) <= 1e-13 * dtb
    assert fa == fb and any(fa)
    assert block_rel_err(ua.conserved_u, ub.conserved_u) <= 1e-12
    assert_scalars(ua.scalars, ub.scalars, 1e-11, ub.time)
    # changing a state in place invalidates the step queued from it
    U = ua.conserved_u
    U[:, 0] *= 1.01
    ua.conserved_u = U
    ub.conserved_u = U
    a.next_solution(ua)
    b.next_solution(ub)
    assert block_rel_err(ua.conserved_u, ub.conserved_u) <= 1e-12
