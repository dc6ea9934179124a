#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native assembly path.

Metric (BASELINE.json): assembled tets/s (and DOF/s) + fraction of the HBM roofline.
Workload at N=1: BASELINE configs[1] = "P2 anisotropic diffusion (full 3x3 tensor coeff) on 10M-tet
synthetic cube mesh": cube 119^3 hexes x 6 = 10,110,954 tets, Operator<GRAD,FemFix<FEM_P2>> squared with a
symmetric per-tet K(x), order-2 rule (q=4), plus the load vector; structural CSR pattern pre-built (timed
separately as pattern_build_ms).  One step = one Assemble (matrix values + rhs) over the whole mesh.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n HEXES_PER_AXIS] [--impl ours|reference]

N>1 (torchrun, one rank per GPU): the cube is split into box blocks like GenerateParallelepiped
(utils/mesh_utils.cpp:67-108); see DESIGN.md section "multi-GPU".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as entry  # noqa: E402

METRIC = "assembled tets/sec (P2 anisotropic diffusion, FP64, CSR values + rhs)"
UNIT = "tets/s"


def sym_K(xc):
    """SPD K(x) per tet = [[2+x, 1/2, 0],[1/2, 1, -1/4],[0, -1/4, 3]] at the centroid"""
    n = xc.shape[0]
    K = np.zeros((n, 9))
    K[:, 0] = 2 + xc[:, 0]; K[:, 4] = 1; K[:, 8] = 3
    K[:, 1] = K[:, 3] = 0.5
    K[:, 5] = K[:, 7] = -0.25
    return K


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "power_w_max": max(float(r[2]) for r in self.rows),
                "samples": len(self.rows), "reasons": reasons}


def cpu_reference_sample(n_sample, steps=1, warmup=0):
    """The reference's CPU implementation of the path on a bounded sample of the same workload:
    unmodified reference fem3Dtet (oracle/_ref) when it was built in the container, else the C port;
    restated Assembler scatter; all host threads.  Returns (tets/s, info)."""
    O, M = entry.load_oracle()
    cores = os.cpu_count() or 1
    co, te, _ = M.cube_mesh(n_sample, n_sample, n_sample)
    dm = M.DofMap(te, [(O.P2, 1)], nnode=co.shape[0])
    K = sym_K(co[te].mean(axis=1))
    prob = M.Problem([(O.P2, 1)],
                     [dict(trial=0, test=0, opA=O.GRAD, opB=O.GRAD, order=2, ttype=O.T_SYMMETRIC, layout=O.L_PER_TET, D=K)],
                     [dict(test=0, opB=O.IDEN, order=2, ttype=O.T_NULL, layout=O.L_CONST)])
    rc0, cc0 = dm.codes(None)
    rp, ci = M.template_pattern(rc0, cc0, 0, dm.nrows)
    val, rhs = np.zeros(ci.size), np.zeros(dm.nrows)
    kind = "reference" if O.have_ref() else "port"
    times = []
    for it in range(warmup + steps):
        val[:] = 0; rhs[:] = 0
        t0 = time.perf_counter()
        if kind == "reference":
            O.ref_assemble_csr(prob, co, te, rc0, 0, rp, ci, val, rhs, nthreads=cores)
        else:
            XY = co[te].transpose(1, 0, 2)
            A, F = prob.element_matrices(XY, idx=np.arange(te.shape[0]))
            O.scatter_csr(rc0, cc0, A, F, 0, rp, ci, val, rhs)
            cores = 1
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    info = {"kind": kind, "cores": cores, "ntet_sample": int(te.shape[0]), "ms_per_step": dt * 1e3,
            "sample": "cube %d^3 x6 = %d tets of the same P2 anisotropic problem (element matrices by the %s, restated "
                      "Assembler scatter into a pre-built CSR, %d std::threads over cell ranges); the true INMOST scatter is slower"
                      % (n_sample, te.shape[0], "unmodified reference fem3Dtet" if kind == "reference" else "C port of fem3Dtet", cores)
                      + "; mean of %d timed passes" % steps}
    return te.shape[0] / dt, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_s = args.ref_n
    v, info = cpu_reference_sample(n_s, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": info["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": config_dict(args.n, {"timed_sample_hexes_per_axis": n_s, "timed_sample_ntet": info["ntet_sample"],
                                           "note": "the reference arm times a bounded sample (cube %d^3 x 6 = %d tets) of the named workload on the host "
                                                   "cores; value is its per-tet rate, the full %d^3 mesh is NOT assembled on the CPU" % (n_s, info["ntet_sample"], args.n)}),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def config_dict(n, extra):
    d = {"workload": "C2: P2 anisotropic diffusion, symmetric per-tet K(x) 3x3, order-2 rule (q=4), cube %d^3 x 6 tets, "
                     "CSR values + rhs, structural pattern pre-built" % n,
         "hexes_per_axis": n, "ntet": 6 * n ** 3, "l2_policy": "inputs larger than L2 (staging+CSR working set >> 126 MB), no explicit flush"}
    if extra:
        d.update(extra)
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", "--hexes-per-axis", dest="n", type=int, default=119,
                    help="hexes per axis per GPU block (119 -> 10,110,954 tets); under torchrun use the long spelling (--n is ambiguous for its parser)")
    ap.add_argument("--global-n", type=int, default=0, help="strong scaling: fix the GLOBAL mesh to this many hexes per axis (default: weak scaling, --n per GPU)")
    ap.add_argument("--strong-n", type=int, default=0, help="N>1: also measure STRONG scaling of C2 on a fixed global mesh of this many hexes per axis "
                    "(256 -> 100,663,296 tets, the north-star mesh; needs N >= 4), reported under the key strong_scaling; 0 (default) = skip: "
                    "the contract line is weak scaling")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-n", type=int, default=64, help="hexes per axis of the CPU sample (64 -> 1,572,864 tets)")
    ap.add_argument("--ref-reps", type=int, default=12, help="timed passes of the CPU sample in the cpu_baseline leg (about 10-20 core-seconds)")
    ap.add_argument("--config", default="c2", choices=["c2", "c4", "c5"],
                    help="multi-GPU runs only: c4 = FemVec<3,P2> elasticity, c5 = Taylor-Hood Stokes (single GPU: bench_configs.py); the contract line is c2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the oracle comparison of small cubes that precedes the timed region")
    ap.add_argument("--no-secondary", action="store_true", help="skip the quick C1/C3/C4/C5 measurements")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the assembly path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import datetime
        # a rank that dies must not leave the others waiting for the default 10 minutes of the NCCL watchdog
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))
    pkg = entry.load_package()
    if world > 1:
        import bench_multi
        return bench_multi.run(args, pkg, rank, world, local_rank)
    if args.config != "c2":
        raise SystemExit("--config c4/c5 is the multi-GPU leg; on one GPU run bench_configs.py --configs c4,c5")

    stream = torch.cuda.current_stream()
    ctx = pkg.Context(local_rank, stream.cuda_stream)
    n = args.global_n or args.n
    t0 = time.perf_counter()
    ctx.mesh_cube(n, n, n)
    ctx.dofmap_natural([(pkg.P2, 1)])
    ctx.sync()
    t1 = time.perf_counter()
    nnz = ctx.pattern_build()
    ctx.sync()
    t2 = time.perf_counter()
    nnode, ntet = ctx.mesh_sizes()
    _, _, _, nrows, _ = ctx.dofmap_info()
    coords, tets = ctx.mesh_get()
    K_host = torch.from_numpy(sym_K(coords[tets].mean(axis=1))).pin_memory()
    del coords, tets
    K_dev = K_host.cuda()
    val_dev = torch.zeros(nnz, dtype=torch.float64, device="cuda")
    rhs_dev = torch.zeros(nrows, dtype=torch.float64, device="cuda")
    mk = lambda K: ([pkg.make_form(pkg.GRAD, pkg.P2, 1, pkg.GRAD, pkg.P2, 1, 2, pkg.TENSOR_SYMMETRIC, pkg.COEF_PER_TET, K)],
                    [pkg.make_form(pkg.IDEN, pkg.P0, 1, pkg.IDEN, pkg.P2, 1, 2, pkg.TENSOR_NULL, pkg.COEF_CONST)])
    forms_d, rhsf_d = mk(K_dev)
    forms_h, rhsf_h = mk(K_host)

    # ---- device-resident arm: inputs already in HBM
    for _ in range(args.warmup):
        assert ctx.assemble(forms_d, rhsf_d, val_dev, rhs_dev) == 0
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    el_ms = ga_ms = 0.0
    ev0.record(stream)
    for _ in range(args.steps):
        assert ctx.assemble(forms_d, rhsf_d, val_dev, rhs_dev) == 0
        t = ctx.last_times()
        el_ms += t["element_ms"]; ga_ms += t["gather_ms"]; fused = t["fused_path"]; names = (t["element_kernel"], t["gather_kernel"])
    ev1.record(stream)
    torch.cuda.synchronize()
    launches = ctx.launch_count()
    ms_dev = ev0.elapsed_time(ev1) / args.steps
    el_ms /= args.steps; ga_ms /= args.steps
    checksum = float(val_dev.sum().item())

    # ---- end-to-end arm: HOST buffers through the C ABI, H2D of the coefficient + D2H of values/rhs every step
    val_host = torch.zeros(nnz, dtype=torch.float64).pin_memory()
    rhs_host = torch.zeros(nrows, dtype=torch.float64).pin_memory()
    e2e_steps = max(2, min(args.steps, 5))
    assert ctx.assemble(forms_h, rhsf_h, val_host, rhs_host) == 0
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(e2e_steps):
        assert ctx.assemble(forms_h, rhsf_h, val_host, rhs_host) == 0
    ev1.record(stream)
    torch.cuda.synchronize()
    ms_e2e = ev0.elapsed_time(ev1) / e2e_steps
    sampler.stop_flag = True
    sampler.join(timeout=2)
    assert abs(float(val_host.sum().item()) - checksum) <= 1e-9 * abs(checksum) + 1e-9

    # ---- roofline of the dominant kernel (algorithmic bytes: every input read once, every output written once)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    alg_bytes = 4 * 10 * ntet + 24 * nnode + 72 * ntet + 8 * nnz + 8 * nrows
    dom_name, dom_ms = (names[0], el_ms) if el_ms >= ga_ms else (names[1], ga_ms)
    traffic, traffic_src = None, None   # measured under ncu (never inside this run): per launch, stamped with kernel + commit
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom_name, {})
        traffic = prof.get("dram_bytes_per_launch")
        traffic_src = "%s @ %s, %s" % (prof.get("kernel"), prof.get("git_sha"), prof.get("report")) if traffic else None
    except Exception:
        pass
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    step_gbs = alg_bytes / (ms_dev * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": traffic,
                "traffic_source": traffic_src, "kernel": dom_name, "kernel_ms": dom_ms, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "step": {"achieved": step_gbs, "frac": step_gbs / peak_gbs, "element_ms": el_ms, "gather_ms": ga_ms, "path": "fused tensor-representation" if fused else "generic staged",
                         "note": "whole step = element kernels + gather; frac of the step is the honest end figure"}}

    line = {"metric": METRIC, "value": ntet / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong" if args.global_n else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(n, {"nnode": nnode, "nrows": nrows, "nnz": nnz, "mesh_dofmap_ms": (t1 - t0) * 1e3, "pattern_build_ms": (t2 - t1) * 1e3}),
            "dof_per_s": nrows / (ms_dev * 1e-3),
            "e2e": {"value": ntet / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "steps": e2e_steps,
                    "h2d_bytes_per_step": int(K_host.numel() * 8), "d2h_bytes_per_step": int((nnz + nrows) * 8)},
            "gpu_launches": int(launches), "roofline": roofline, "clocks": sampler.summary()}
    if not args.no_secondary:
        # the other BASELINE.json configs at sizes that keep the default run short (parity-test cases, not bench lines): evidence
        # of which product path serves them and at what fraction of their own HBM roofline
        import bench_configs
        sec = {}
        for name, nn in (("c1", 96), ("c3", 40), ("c4", 48), ("c5", 48)):
            try:
                r = bench_configs.run_config(pkg, name, nn, 3, stream, peak_gbs)
                sec[name] = {k: r[k] for k in ("ntet", "nnz", "ms_per_assemble", "tets_per_s", "dof_per_s", "kernels", "path", "hbm_frac")}
            except Exception as exc:  # noqa: BLE001 -- a secondary measurement must not void the headline line
                sec[name] = {"error": str(exc)[:200]}
        line["secondary_configs"] = sec
    if not args.no_cpu_baseline:
        v, info = cpu_reference_sample(min(args.ref_n, 48), steps=args.ref_reps, warmup=1)   # bounded: ~10-20 core-seconds
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]}
    print(json.dumps(line))
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
