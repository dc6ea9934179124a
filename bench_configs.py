#!/usr/bin/env python3
"""Secondary measurements: throughput of every BASELINE.json config (C1..C5) on one GPU at a size that fits a quick run.
Not the contract bench (bench.py is): one JSON line per config with tets/s, DOF/s, the path that ran and the fraction of the
HBM roofline on the algorithmic bytes of SURVEY.md section 8d.

  python bench_configs.py [--n HEXES_PER_AXIS] [--steps K] [--configs c1,c2,...]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
import golden_cases as gc  # noqa: E402
import problems  # noqa: E402


def run_config(pkg, name, n, steps, stream, peak):
    """assemble config `name` on an n^3 x 6 cube on the current GPU; returns the measurement as a dict"""
    import torch
    variables = {"c1": [(gc.P1, 1)], "c2": [(gc.P2, 1)], "c3": [(gc.P3, 1)], "c4": [(gc.P2, 3)], "c5": [(gc.P2, 3), (gc.P1, 1)]}[name]
    ctx = pkg.Context(torch.cuda.current_device(), stream.cuda_stream)
    ctx.mesh_cube(n, n, n)
    ctx.dofmap_natural(variables)
    t0 = time.perf_counter()
    nnz = ctx.pattern_build()
    ctx.sync()
    t_pat = (time.perf_counter() - t0) * 1e3
    coords, tets = ctx.mesh_get()
    nnode, ntet = coords.shape[0], tets.shape[0]
    _, _, _, nrows, _ = ctx.dofmap_info()
    if name == "c1":
        _, forms, rhsf, _ = problems.c1_p1_diffusion(pkg, None, coords, tets)
        coef_bytes = 0
    elif name == "c2":
        _, forms, rhsf, _ = problems.c2_p2_aniso(pkg, None, coords, tets)
        coef_bytes = 72
    elif name == "c3":
        XY = coords[tets].transpose(1, 0, 2)
        xyg4, xyg6 = ctx.quad_points(4, XY), ctx.quad_points(6, XY)
        _, forms, rhsf, _ = problems.c3_p3_react_diff(pkg, None, coords, tets, xyg4, xyg6)
        coef_bytes = 8 * (14 + 24)
    elif name == "c4":
        _, forms, rhsf, _ = problems.c4_p2_elasticity(pkg, None, coords, tets)
        coef_bytes = 0
    else:
        _, forms, rhsf, _ = problems.c5_stokes(pkg, None, coords, tets)
        coef_bytes = 0
    # coefficients to the device once (device-resident arm)
    dev = []
    for f in forms + rhsf:
        if f._keep is not None:
            t = torch.from_numpy(np.ascontiguousarray(f._keep)).cuda()
            dev.append(t)
            f.D, f.coef_space, f._keep = t.data_ptr(), pkg.DEVICE, t
    val = torch.zeros(nnz, dtype=torch.float64, device="cuda")
    rhs = torch.zeros(nrows, dtype=torch.float64, device="cuda")
    for _ in range(2):
        assert ctx.assemble(forms, rhsf, val, rhs) == 0
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        assert ctx.assemble(forms, rhsf, val, rhs) == 0
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    t = ctx.last_times()
    nloc = sum(gc.NF[f] * v for f, v in variables)
    alg = 4 * nloc * ntet + 24 * nnode + coef_bytes * ntet + 8 * nnz + 8 * nrows
    out = {"config": name, "hexes_per_axis": n, "ntet": ntet, "nrows": nrows, "nnz": nnz, "ms_per_assemble": ms,
           "tets_per_s": ntet / (ms * 1e-3), "dof_per_s": nrows / (ms * 1e-3), "element_ms": t["element_ms"],
           "gather_ms": t["gather_ms"], "kernels": [t["element_kernel"], t["gather_kernel"]],
           "path": "fused tensor-representation" if t["fused_path"] else "generic staged",
           "algorithmic_bytes_per_tet": alg / ntet, "hbm_frac": alg / (ms * 1e-3) / 1e9 / peak, "pattern_build_ms": t_pat}
    if name == "c3":
        # contraction-bound case: useful FLOPs of fem3Dtet for the full 20 x 20 matrix (stiffness q = 14: 3 x 400 + 180 FMA per
        # point, mass q = 24: 400 FMA per point) against the FP64 peak measured with tools/fp64_peak.cu on this pool's B200
        flop = 2.0 * (14 * (3 * 400 + 180) + 24 * 400)
        fp64_peak = 36.74e12
        try:
            fp64_peak = json.load(open(os.path.join(ROOT, "profiles", "r02", "r02c_fp64_peak.json")))["dfma_tflops"] * 1e12
        except Exception:
            pass
        out["fp64"] = {"useful_flop_per_tet": flop, "peak_tflops": fp64_peak / 1e12, "peak_source": "tools/fp64_peak.cu (profiles/r02/r02c_fp64_peak.json)",
                       "frac_element_stage": flop * ntet / (t["element_ms"] * 1e-3) / fp64_peak, "frac_assembly": flop * ntet / (ms * 1e-3) / fp64_peak}
    ctx.close()
    del val, rhs, dev
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=48)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--configs", default="c1,c2,c3,c4,c5")
    args = ap.parse_args()
    import torch
    pkg = entry.load_package()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    stream = torch.cuda.current_stream()
    for name in args.configs.split(","):
        n = args.n if name in ("c1", "c2") else max(8, args.n // 2)
        print(json.dumps(run_config(pkg, name, n, args.steps, stream, peak)))


if __name__ == "__main__":
    main()
