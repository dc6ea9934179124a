"""N>1 leg of bench.py: one rank per GPU (torchrun), the cube split into box blocks like GenerateParallelepiped
(utils/mesh_utils.cpp:67-108), every GPU assembles its own elements and the contributions to interface rows are
exchanged over NCCL (owner-computes + exchange, inmost-fem_b200/parallel.py).  Weak scaling: the per-GPU block is
the N=1 workload (n^3 hexes), so the global mesh grows with the number of GPUs."""
import importlib
import json
import sys
import os
import time

import torch
import torch.distributed as dist

import bench


PARITY_CASES = (("p2", (12, 10, 8)), ("taylor_hood", (6, 5, 4)), ("p3", (5, 4, 3)))


def parity_check(pkg, par, rank, world, local_rank):
    """Correctness of the path the timed region measures, on the same process group: small cubes (P2, Taylor-Hood, P3) are
    assembled by all ranks and every rank's owned rows are compared with the CPU oracle's per-rank matrices
    (oracle/asm_oracle.py, used here as the checker only): numbering and CSR pattern bit-exact, values / rhs relative to the
    largest entry of the row <= 1e-12.  Returns the dict for the JSON line; `ok` is the AND over all ranks and cases."""
    import numpy as np
    import golden_cases as gc
    import problems
    O, M = bench.entry.load_oracle()
    checked, worst, ok = [], 0.0, True
    for name, dims in PARITY_CASES:
        variables = {"p2": [(gc.P2, 1)], "taylor_hood": [(gc.P2, 3), (gc.P1, 1)], "p3": [(gc.P3, 1)]}[name]
        ctx = pkg.Context(local_rank, torch.cuda.current_stream().cuda_stream)
        case_ok, err, agreed = True, 0.0, False
        try:
            da = par.DistributedAssembler(ctx, dims, variables)
            co, te, cr = M.cube_mesh(*dims, nranks=world)
            dm = M.DofMap(te, variables, cr, world, nnode=co.shape[0])
            mine = np.nonzero(cr == rank)[0]
            case_ok &= bool(np.array_equal(da.numbering.elem2dof.cpu().numpy(), dm.elem2dof[mine]))
            xc = co[te].mean(axis=1)
            if name == "taylor_hood":
                _, forms, rhsf, prob = problems.c5_stokes(pkg, M, co, te)
            else:
                fem = variables[0][0]
                mats = [(0, 0, gc.GRAD, gc.GRAD, 2, gc.T_SYMMETRIC, gc.L_PER_TET, problems.sym_K(xc), 1.0)]
                rhss = [(0, gc.IDEN, 2, gc.T_NULL, gc.L_CONST, None, 1.0)]
                _, _, _, prob = problems._mk(pkg, M, variables, mats, rhss)
                K_loc = torch.from_numpy(problems.sym_K(xc[mine])).cuda()
                forms = [pkg.make_form(gc.GRAD, fem, 1, gc.GRAD, fem, 1, 2, gc.T_SYMMETRIC, gc.L_PER_TET, K_loc)]
                rhsf = [pkg.make_form(gc.IDEN, gc.P0, 1, gc.IDEN, fem, 1, 2, gc.T_NULL, gc.L_CONST)]
            rp_o, ci_o, v_o, r_o, _ = M.assemble(prob, co, te, dm, rank=rank)
            case_ok &= bool(np.array_equal(da.rowptr.cpu().numpy(), rp_o) and np.array_equal(da.colind.cpu().numpy(), ci_o))
            # the assembly contains an exchange: every rank enters it or none does (a rank that failed a check on its own must
            # not leave its peers waiting in the exchange)
            agree = torch.tensor([1.0 if case_ok else 0.0], device="cuda")
            dist.all_reduce(agree, op=dist.ReduceOp.MIN)
            agreed = True
            if case_ok and agree.item() == 1.0:
                assert da.assemble(forms, rhsf) == 0
                torch.cuda.synchronize()
                val = da.val[:da.plan.nnz_own].cpu().numpy()
                rhs = da.rhs[:da.plan.n_own].cpu().numpy()
                rowmax = np.maximum.reduceat(np.abs(v_o), rp_o[:-1])
                err = float((np.abs(val - v_o) / np.repeat(rowmax, np.diff(rp_o))).max())
                err = max(err, float(np.abs(rhs - r_o).max() / np.abs(r_o).max()))
                case_ok &= err <= 1e-12
        except Exception as exc:  # noqa: BLE001 -- a failing case is reported (and fails the run), it must not hang the other ranks
            case_ok, err = False, float("inf")
            print("parity case %s raised on rank %d: %r" % (name, rank, exc), file=sys.stderr, flush=True)
            if not agreed:   # keep the collective sequence of the other ranks matched
                agree = torch.tensor([0.0], device="cuda")
                dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        finally:
            ctx.close()
        t = torch.tensor([0.0 if case_ok else 1.0, err if err == err and err != float("inf") else 1e300], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok &= t[0].item() == 0.0
        worst = max(worst, t[1].item())
        checked.append({"case": name, "hexes": list(dims), "ok": t[0].item() == 0.0, "max_rel_err": t[1].item()})
    return {"ok": ok, "checked": checked, "max_rel_err": worst, "tolerance": 1e-12, "pattern": "bit-exact vs oracle/asm_oracle.py per rank",
            "ranks": world}


def strong_leg(args, pkg, par, rank, world, local_rank, global_n):
    """Second measurement of the contract line for N > 1: STRONG scaling of C2 on the north-star mesh (global_n^3 hexes x 6 tets
    = 100.66 M tets at 256), blocks shrinking with the number of GPUs.  Device-resident steps only (CUDA events, max over ranks)."""
    dims = (global_n,) * 3
    ppa = par.proc_grid(world, dims)
    stream = torch.cuda.current_stream()
    ctx = pkg.Context(local_rank, stream.cuda_stream)
    t0 = time.perf_counter()
    da = par.DistributedAssembler(ctx, dims, [(pkg.P2, 1)])
    torch.cuda.synchronize()
    setup_ms = (time.perf_counter() - t0) * 1e3
    ntet = da.ntet
    xc = da.coords[da.tets.long()].mean(dim=1)
    K_dev = torch.zeros((ntet, 9), dtype=torch.float64, device="cuda")
    K_dev[:, 0] = 2 + xc[:, 0]; K_dev[:, 4] = 1; K_dev[:, 8] = 3
    K_dev[:, 1] = K_dev[:, 3] = 0.5
    K_dev[:, 5] = K_dev[:, 7] = -0.25
    del xc
    forms = [pkg.make_form(pkg.GRAD, pkg.P2, 1, pkg.GRAD, pkg.P2, 1, 2, pkg.TENSOR_SYMMETRIC, pkg.COEF_PER_TET, K_dev)]
    rhsf = [pkg.make_form(pkg.IDEN, pkg.P0, 1, pkg.IDEN, pkg.P2, 1, 2, pkg.TENSOR_NULL, pkg.COEF_CONST)]
    for _ in range(args.warmup):
        assert da.assemble(forms, rhsf) == 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(args.steps):
        assert da.assemble(forms, rhsf) == 0
    ev1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1) / args.steps], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    tot = torch.tensor([ntet, da.plan.n_own, da.plan.nnz_own], dtype=torch.int64, device="cuda")
    dist.all_reduce(tot)
    ntet_all, nrows_all, nnz_all = [int(v) for v in tot.tolist()]
    nn_all = (global_n + 1) ** 3
    alg_bytes = 40 * ntet_all + 24 * nn_all + 72 * ntet_all + 8 * nnz_all + 8 * nrows_all
    out = {"scaling": "strong", "global_hexes": list(dims), "proc_grid": ppa, "ntet": ntet_all, "nrows": nrows_all, "nnz": nnz_all,
           "ms_per_step": ms.item(), "value": ntet_all / (ms.item() * 1e-3), "unit": bench.UNIT, "dof_per_s": nrows_all / (ms.item() * 1e-3),
           "hbm_frac_per_gpu": alg_bytes / world / (ms.item() * 1e-3) / 1e9 / bench_peak(), "setup_ms_rank0": setup_ms, "steps": args.steps}
    ctx.close()
    del da, K_dev
    torch.cuda.empty_cache()
    return out


def bench_peak():
    try:
        return json.load(open(os.path.join(bench.ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)
    except Exception:
        return 6650.0


def run(args, pkg, rank, world, local_rank):
    par = importlib.import_module("inmost_fem_b200.parallel")
    parity = None
    if not getattr(args, "no_parity", False):
        parity = parity_check(pkg, par, rank, world, local_rank)
        # the case that exercises the timed workload's own space decides whether a number may be printed at all; a failure of
        # another case (e.g. P3 while C2 = P2 is timed) is reported in the line (parity.ok = false) and fails the run at the end
        own_case = "p2" if getattr(args, "config", "c2") == "c2" else "taylor_hood"
        own_ok = all(c["ok"] for c in parity["checked"] if c["case"] == own_case)
        if not own_ok:
            if rank == 0:
                print(json.dumps({"error": "multi-GPU parity check failed", "parity": parity}))
            dist.barrier()
            dist.destroy_process_group()
            return 3
    n = args.n
    if args.global_n:
        # strong scaling: the global mesh is fixed, the blocks shrink with the number of GPUs
        dims = (args.global_n,) * 3
        ppa = par.proc_grid(world, dims)
    else:
        ppa = par.proc_grid(world, (n, n, n))
        dims = (n * ppa[0], n * ppa[1], n * ppa[2])
        assert par.proc_grid(world, dims) == ppa
    stream = torch.cuda.current_stream()
    ctx = pkg.Context(local_rank, stream.cuda_stream)
    t0 = time.perf_counter()
    config = getattr(args, "config", "c2")
    variables = {"c2": [(pkg.P2, 1)], "c4": [(pkg.P2, 3)], "c5": [(pkg.P2, 3), (pkg.P1, 1)]}[config]
    da = par.DistributedAssembler(ctx, dims, variables)
    torch.cuda.synchronize()
    setup_ms = (time.perf_counter() - t0) * 1e3
    ntet = da.ntet
    if config == "c2":
        xc = da.coords[da.tets.long()].mean(dim=1)
        K_dev = torch.zeros((ntet, 9), dtype=torch.float64, device="cuda")
        K_dev[:, 0] = 2 + xc[:, 0]; K_dev[:, 4] = 1; K_dev[:, 8] = 3
        K_dev[:, 1] = K_dev[:, 3] = 0.5
        K_dev[:, 5] = K_dev[:, 7] = -0.25
        K_host = K_dev.cpu().pin_memory()
        mk = lambda K: ([pkg.make_form(pkg.GRAD, pkg.P2, 1, pkg.GRAD, pkg.P2, 1, 2, pkg.TENSOR_SYMMETRIC, pkg.COEF_PER_TET, K)],
                        [pkg.make_form(pkg.IDEN, pkg.P0, 1, pkg.IDEN, pkg.P2, 1, 2, pkg.TENSOR_NULL, pkg.COEF_CONST)])
        forms_d, rhsf_d = mk(K_dev)
        forms_h, rhsf_h = mk(K_host)
        h2d_bytes = int(K_host.numel() * 8)
    else:
        # C4 (FemVec<3,P2> elasticity, constant 9x9 tensor) / C5 (Taylor-Hood Stokes): constant coefficients, nothing per tet to ship
        import problems
        _, forms_d, rhsf_d, _ = (problems.c4_p2_elasticity if config == "c4" else problems.c5_stokes)(pkg, None, None, None)
        forms_h, rhsf_h = forms_d, rhsf_d
        h2d_bytes = 0
    for _ in range(args.warmup):
        assert da.assemble(forms_d, rhsf_d) == 0
    sampler = bench.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    ctx.launch_count(reset=True)
    ev0.record(stream)
    for _ in range(args.steps):
        assert da.assemble(forms_d, rhsf_d) == 0
    ev1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    launches = ctx.launch_count()
    ms = torch.tensor([ev0.elapsed_time(ev1) / args.steps], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # end to end: host coefficient in, owned values + rhs out, every step
    nnz_own, n_own = da.plan.nnz_own, da.plan.n_own
    val_host = torch.zeros(nnz_own, dtype=torch.float64).pin_memory()
    rhs_host = torch.zeros(n_own, dtype=torch.float64).pin_memory()
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_step():
        assert da.assemble(forms_h, rhsf_h) == 0
        val_host.copy_(da.val[:nnz_own], non_blocking=True)
        rhs_host.copy_(da.rhs[:n_own], non_blocking=True)
    e2e_step()
    dist.barrier()
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(e2e_steps):
        e2e_step()
    ev1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    ms_e2e = torch.tensor([ev0.elapsed_time(ev1) / e2e_steps], device="cuda")
    dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    tot = torch.tensor([ntet, n_own, nnz_own, da.plan.n_for, sum(da.plan.send_nnz)], dtype=torch.int64, device="cuda")
    dist.all_reduce(tot)
    times = ctx.last_times()
    c_exch = bool(getattr(da, "c_exchange", False))
    strong = None
    if config == "c2" and not args.global_n and getattr(args, "strong_n", 0) > 0:
        if rank == 0:
            sampler.stop_flag = True
            sampler.join(timeout=2)
        ctx.close()
        del da, K_dev, K_host, val_host, rhs_host, forms_d, rhsf_d, forms_h, rhsf_h, mk, xc
        torch.cuda.empty_cache()
        # guard: the setup (pattern union in torch: several int64 arrays per non-zero) peaks near 3.8 kB per local tet -- measured:
        # 50.3 M tets per GPU (N = 2) runs out of the 180 GB, 12.6 M (N = 8) fits.  Every rank takes the same decision.
        ntet_loc = 6 * args.strong_n ** 3 // world
        free_b, _ = torch.cuda.mem_get_info()
        fits = torch.tensor([1.0 if 3800.0 * ntet_loc < 0.8 * free_b else 0.0], device="cuda")
        dist.all_reduce(fits, op=dist.ReduceOp.MIN)
        if fits.item() == 1.0:
            strong = strong_leg(args, pkg, par, rank, world, local_rank, args.strong_n)
        else:
            strong = {"scaling": "strong", "global_hexes": [args.strong_n] * 3, "skipped": "%.1f M tets per GPU do not fit the setup's peak memory "
                      "(about 3.8 kB per tet in the torch pattern union); run with more GPUs" % (ntet_loc / 1e6)}
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
        ntet_all, nrows_all, nnz_all, nfor_all, nsend_all = [int(v) for v in tot.tolist()]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(bench.ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_gbs = peaks.get("hbm_gbs", 6650.0)
        nn_all = (dims[0] + 1) * (dims[1] + 1) * (dims[2] + 1)
        nloc = sum({pkg.P1: 4, pkg.P2: 10}[f] * v for f, v in variables)
        alg_bytes = 4 * nloc * ntet_all + 24 * nn_all + (72 * ntet_all if config == "c2" else 0) + 8 * nnz_all + 8 * nrows_all
        gbs = alg_bytes / (ms.item() * 1e-3) / 1e9
        metric = bench.METRIC if config == "c2" else "assembled tets/sec (%s, FP64, CSR values + rhs)" % (
            "C4: FemVec<3,P2> linear elasticity, constant 9x9 tensor" if config == "c4" else "C5: Taylor-Hood P2^3 x P1 Stokes")
        line = {"metric": metric, "value": ntet_all / (ms.item() * 1e-3), "unit": bench.UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms.item(), "higher_is_better": True, "scaling": "strong" if args.global_n else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": bench.config_dict(n, {"global_hexes": list(dims), "proc_grid": ppa, "ntet": ntet_all, "nrows": nrows_all, "nnz": nnz_all,
                                                "interface_rows_sent": nfor_all, "interface_values_sent_per_step": nsend_all,
                                                "exchange": ("grouped ncclSend/ncclRecv issued by the library (afb_assemble_distributed) + additions in rank order" if c_exch else "NCCL all_to_all_single of packed FP64 interface contributions + afb_halo_add in rank order"),
                                                "setup_ms_rank0": setup_ms}),
                "dof_per_s": nrows_all / (ms.item() * 1e-3),
                "e2e": {"value": ntet_all / (ms_e2e.item() * 1e-3), "unit": bench.UNIT, "ms_per_step": ms_e2e.item(), "steps": e2e_steps,
                        "h2d_bytes_per_step": h2d_bytes * world, "d2h_bytes_per_step": int((nnz_all + nrows_all) * 8)},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": gbs / world, "peak": peak_gbs, "unit": "GB/s", "frac": gbs / world / peak_gbs, "traffic": None,
                             "kernel": "whole step per GPU (%s + %s + exchange)" % (times["element_kernel"], times["gather_kernel"]),
                             "algorithmic_bytes_per_launch": alg_bytes // world},
                "clocks": sampler.summary(), "parity": parity}
        if strong is not None:
            # second key: strong scaling on the north-star mesh (the contract line above is weak scaling: fixed per-GPU block)
            line["strong_scaling"] = strong
        if config != "c2":
            line["config"]["workload"] = metric + ", per-GPU block %d^3 hexes x 6 tets, structural pattern pre-built" % n
        print(json.dumps(line))
    if strong is None:
        ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0 if (parity is None or parity["ok"]) else 3
