"""The reference-side binding of INTEGRATION.md, compiled and run (oracle/shim/subprog_binary_b200.cpp).

oracle/_ref/mara_ref       the reference's own `binary` translation units + oracle/ref_harness.cpp
oracle/_ref/mara_ref_b200  the same, but subprog_binary_scheme.cpp replaced by the shim, which implements
                           binary::advance / maximum_timestep / set_scheme_globals / recover_primitive and
                           solution_t::operator+ / operator* on the C ABI of libmara3_b200.so

Both run the reference's own next_solution (dt rule, RK2 through operator+ / operator*, try / catch retry); the only
difference is who computes a stage.  Their states must agree to the per-cell tolerance of the path (1e-12) and their dt
histories to 1e-13 -- which is the drop-in claim, tested through the reference's own data structures."""
import os
import subprocess
import numpy as np
import pytest
from conftest import block_rel_err, block_rel_err_q
from oracle_util import REF_BIN, ORACLE_DIR, read_dump, dump_scalars

SHIM_BIN = os.path.join(ORACLE_DIR, "_ref", "mara_ref_b200")
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(SHIM_BIN)),
                                 reason="oracle/_ref/mara_ref and mara_ref_b200 are built where /root/reference exists (make -C oracle ref shim)")]


def run(binary, cfg, steps, dump):
    cmd = [binary, "--steps", str(steps), "--dump", dump, "--dump-steps", str(steps)] + [f"{k}={v}" for k, v in cfg.items()]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (cmd, r.stdout[-2000:], r.stderr[-2000:])
    return read_dump(dump), r.stdout


@pytest.mark.parametrize("cfg,steps", [
    (dict(depth=3, block_size=32, domain_radius=6.0, focus_factor=1e3, threaded=4), 3),     # uniform 256^2: stage_tma
    (dict(depth=4, block_size=32, threaded=4), 3),                                          # nested: jump blocks, flux correction
    (dict(depth=2, block_size=64, threaded=4), 25),                                         # config 1: runs into the safe-mode retries (steps 23-25)
    (dict(depth=3, block_size=16, conserve_linear_p=0, fixed_dt=1, threaded=4), 3),         # advance_q
])
def test_reference_run_loop_on_the_b200_library_matches_the_reference(cfg, steps, tmp_path):
    ref, out_ref = run(REF_BIN, cfg, steps, str(tmp_path / "ref.m3bd"))
    got, out_got = run(SHIM_BIN, cfg, steps, str(tmp_path / "b200.m3bd"))
    key = f"step{steps}/conserved_u"
    rel_err = block_rel_err if cfg.get("conserve_linear_p", 1) else block_rel_err_q
    # three steps of per-stage rounding differences (FMA contraction, reduction order): the per-step bound is 1e-12
    assert rel_err(got[key], ref[key]) <= 1e-12 * max(1, steps // 3)
    assert np.allclose(got["dt_history"], ref["dt_history"], rtol=1e-12, atol=0.0)
    a, b = dump_scalars(got, f"step{steps}/"), dump_scalars(ref, f"step{steps}/")
    assert a[0] == pytest.approx(b[0], rel=1e-12) and a[1] == b[1] and a[2] == b[2]       # time, iteration (rational)
    # the retry fires on the same steps (the harness prints the fallback count)
    fb = lambda text: [t for t in text.split() if t.startswith("fallbacks=")]
    assert fb(out_got) == fb(out_ref)
