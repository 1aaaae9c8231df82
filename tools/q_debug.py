"""Development aid: where the conserved_q (advance_q) path differs most from the oracle on a nested tree, strip kernel against any-tree kernels."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mara3_b200 as m3
from oracle_util import OracleMesh, OracleSolution
cfg = dict(depth=4, block_size=32, conserve_linear_p=0, fixed_dt=1)
o = OracleSolution(OracleMesh(cfg))
dt = 0.4 * o.maximum_timestep()
o1, st = o.advance(dt)
for general in (False, True):
    s = m3.Solver(cfg, general_only=general); u = s.create_solution()
    g1 = s.advance(u, dt)
    a, b = g1.conserved_u, o1.conserved_u
    scale = np.abs(b).max(axis=(2, 3), keepdims=True)
    rp = np.hypot(b[:, 1], b[:, 2]).max(axis=(1, 2)); scale[:, 1, 0, 0] = rp; scale[:, 2, 0, 0] = rp
    err = np.abs(a - b) / scale
    print("general_only", general, "regular", s.num_regular_blocks, "of", s.num_blocks, "max err", err.max())
    idx = np.argsort(err.ravel())[::-1][:12]
    ti = s.tree_index
    for k in idx:
        blk, f, i, j = np.unravel_index(k, err.shape)
        cc = s.cell_centers[blk][:, i, j]
        print("  block", blk, "level/ij", ti[blk], "field", f, "cell", i, j, "xy", cc, "err %.2e" % err[blk, f, i, j], "val", a[blk, f, i, j], b[blk, f, i, j])
