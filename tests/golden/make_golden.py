"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF
(oracle/_ref/mara_ref = Mara3's own sources compiled in place by oracle/Makefile, parity
flags g++ -O2).  Needs /root/reference (build container only):

    make -C oracle ref && python tests/golden/make_golden.py

Each fixture is an .npz with the config, the mesh arrays the reference built, and the
solution (conserved field + 43 scalars) it reached after the listed steps."""
import json
import os
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_util import run_reference, read_dump, dump_scalars

CASES = {
    # nested default focusing, three levels with jumps (28 leaves)
    "nested_d3_n8": dict(config=dict(depth=3, block_size=8), steps=[1, 4], stages=True),
    # uniform, small
    "uniform_d2_n16": dict(config=dict(depth=2, block_size=16, focus_factor=1e3), steps=[1, 5], stages=True),
    # eccentric unequal-mass live binary with constant-nu viscosity, cutoff, floor, inflow
    "live_ecc_d3_n8": dict(config=dict(depth=3, block_size=8, eccentricity=0.3, mass_ratio=0.5, nu=0.01, alpha_cutoff_radius=1.0,
                                       begin_live_binary=0.0, density_floor=1e-2, mdot=1e-4), steps=[1, 5], stages=False),
    # RK1, fixed dt, axisymmetric sound speed, retrograde disk
    "rk1_axisym_d3_n8": dict(config=dict(depth=3, block_size=8, axisymmetric_cs2=1, counter_rotate=1, rk_order=1, fixed_dt=1,
                                         no_accretion_force=1), steps=[1, 5], stages=False),
    # angular-momentum-conserving variables (advance_q): conserve_linear_p=0 needs fixed_dt=1 in the reference;
    # "conserved_u" then holds conserved_q = (sigma, Sr, Lz)
    "angmom_nested_d3_n8": dict(config=dict(depth=3, block_size=8, conserve_linear_p=0, fixed_dt=1), steps=[1, 4], stages=True),
    "angmom_live_rk1_d2_n16": dict(config=dict(depth=2, block_size=16, domain_radius=6.0, conserve_linear_p=0, fixed_dt=1, rk_order=1,
                                               eccentricity=0.2, mass_ratio=0.5, begin_live_binary=0.0, nu=0.01), steps=[1, 5], stages=False),
    # deeper nested tree with a non power-of-two block size: topology + geometry only
    "mesh_d5_n12": dict(config=dict(depth=5, block_size=12), steps=[], stages=False),
    # default run config: depth=4 block_size=24 (64 leaves): topology + geometry only
    "mesh_default": dict(config=dict(), steps=[], stages=False),
}


def main():
    for name, case in CASES.items():
        with tempfile.TemporaryDirectory() as td:
            f = os.path.join(td, "dump.bin")
            steps = case["steps"]
            out = run_reference(dict(case["config"], threaded=4), steps=max(steps) if steps else 0, dump=f,
                                dump_steps=steps, stages=case["stages"], mesh_only=not steps)
            d = read_dump(f)
        arrays = {"config_json": np.array(json.dumps(case["config"])), "steps": np.array(steps, dtype=np.int64),
                  "tree_index": d["tree_index"]}
        small_mesh = d["tree_index"].shape[0] * d["vertices"].shape[-1] ** 2 < 40000
        if small_mesh:
            for k in ("vertices", "cell_centers", "cell_areas", "buffer_rate_field", "initial_conserved_u"):
                arrays[k] = d[k]
        else:   # keep large meshes small in git: block corners + checksums
            arrays["vertex_corners"] = d["vertices"][:, :, [0, -1], :][:, :, :, [0, -1]]
            for k in ("vertices", "cell_centers", "cell_areas", "buffer_rate_field", "initial_conserved_u"):
                arrays[k + "_sum"] = np.array([d[k].sum(), np.abs(d[k]).sum(), (d[k] ** 2).sum()])
        for k in ("recommended_time_step", "gst_suppr_radius", "density_floor"):
            arrays[k] = d[k]
        if steps:
            arrays["dt_history"] = d["dt_history"]
        for n in steps:
            arrays[f"step{n}_conserved_u"] = d[f"step{n}/conserved_u"]
            arrays[f"step{n}_scalars"] = dump_scalars(d, f"step{n}/")
            if case["stages"]:
                for st in (1, 2):
                    arrays[f"step{n}_stage{st}_conserved_u"] = d[f"step{n}/stage{st}/conserved_u"]
                    arrays[f"step{n}_stage{st}_scalars"] = dump_scalars(d, f"step{n}/stage{st}/")
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(name, out.strip(), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
