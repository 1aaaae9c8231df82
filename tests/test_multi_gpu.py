"""GPU, needs >= 2 devices: the distributed solver (NCCL guard-zone exchange, one process per GPU)
against the single-GPU solver and the oracle.  Skipped on single-GPU boxes; the host-side logic of
the same path is covered on the CPU with gloo in test_multi_rank_cpu.py."""
import os
import subprocess
import sys
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_match_one():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert "MULTI_GPU_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_a_stalled_peer_is_reported_not_waited_for_forever():
    """Every wait on another rank is bounded (bounded_wait_sys): a rank whose peer stops stepping gets M3B_ERROR naming the peer
    after the deadline, instead of spinning on a flag for ever (tools/stalled_peer_check.py)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29519", os.path.join(ROOT, "tools", "stalled_peer_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert "STALLED_PEER_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_two_rank_subprogram_writes_the_same_products(tmp_path):
    """`binary` on two GPUs (tools/run_binary.py under torchrun) against the single-GPU executable: both ranks write
    their own blocks into the one checkpoint / diagnostics file; fields bit-identical, time-series sums to rounding (per-rank partial sums)."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from h5_reader import H5File
    args = ["binary", "depth=4", "block_size=32", "tfinal=0.01", "cpi=0.004", "dfi=0.004", "tsi=0.002"]     # nested tree
    one, two = str(tmp_path / "one"), str(tmp_path / "two")
    r1 = subprocess.run([os.path.join(ROOT, "mara3_b200", "bin", "mara3b")] + args + ["outdir=" + one], capture_output=True, text=True, timeout=600)
    assert r1.returncode == 0, r1.stdout[-2000:] + r1.stderr[-2000:]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29521", os.path.join(ROOT, "tools", "run_binary.py")] + args + ["outdir=" + two]
    r2 = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r2.returncode == 0, r2.stdout[-3000:] + r2.stderr[-3000:]
    steps1 = [l.split(" kzps")[0] for l in r1.stdout.splitlines() if l.startswith("[")]
    steps2 = [l.split(" kzps")[0] for l in r2.stdout.splitlines() if l.startswith("[")]
    assert steps1 == steps2 and len(steps1) > 3                                     # same iteration numbers and times, printed once
    assert sorted(os.listdir(one)) == sorted(os.listdir(two))
    for name in sorted(os.listdir(one)):
        a, b = H5File(os.path.join(one, name)), H5File(os.path.join(two, name))
        if name.startswith("chkpt"):
            assert a.read("/solution/time") == b.read("/solution/time")
            leaves = a.keys("/solution/conserved_u")
            assert leaves == b.keys("/solution/conserved_u") and len(leaves) == 64
            for leaf in leaves:
                assert np.array_equal(a.read("/solution/conserved_u/" + leaf), b.read("/solution/conserved_u/" + leaf))
            sa, sb = a.read("/time_series"), b.read("/time_series")
            assert np.array_equal(sa["time"], sb["time"]) and np.allclose(sa["disk_mass"], sb["disk_mass"], rtol=1e-14, atol=0)
        else:
            for group in ("sigma", "radial_velocity", "phi_velocity", "vertices"):
                for leaf in a.keys("/" + group):
                    assert np.array_equal(a.read(f"/{group}/{leaf}"), b.read(f"/{group}/{leaf}"))
