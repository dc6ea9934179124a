"""GPU test of the multi-GPU path (needs >= 2 GPUs: run with `gpurun --gpus 2`): one process per GPU over NCCL,
owner-computes + interface exchange, compared rank by rank with the oracle's per-rank matrices (pattern bit-exact,
values / rhs 1e-12)."""
import os
import sys
import traceback

import numpy as np
import pytest
import torch

from conftest import ROOT

import golden_cases as gc

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, dims, variables, errq):
    try:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import importlib
        import __graft_entry__ as entry
        import problems
        pkg = entry.load_package()
        O, M = entry.load_oracle()
        par = importlib.import_module("inmost_fem_b200.parallel")
        ctx = pkg.Context(rank, torch.cuda.current_stream().cuda_stream)
        da = par.DistributedAssembler(ctx, dims, variables)
        co, te, cr = M.cube_mesh(*dims, nranks=world)
        dm = M.DofMap(te, variables, cr, world, nnode=co.shape[0])
        mine = np.nonzero(cr == rank)[0]
        assert np.array_equal(da.numbering.elem2dof.cpu().numpy(), dm.elem2dof[mine])
        xc_all = co[te].mean(axis=1)
        if variables[0][1] == 1:
            mats = [(0, 0, gc.GRAD, gc.GRAD, 2, gc.T_SYMMETRIC, gc.L_PER_TET, problems.sym_K(xc_all), 1.0)]
            rhss = [(0, gc.IDEN, 2, gc.T_NULL, gc.L_CONST, None, 1.0)]
            _, _, _, prob = problems._mk(pkg, M, variables, mats, rhss)
            K_loc = torch.from_numpy(problems.sym_K(xc_all[mine])).cuda()
            forms = [pkg.make_form(gc.GRAD, variables[0][0], 1, gc.GRAD, variables[0][0], 1, 2, gc.T_SYMMETRIC, gc.L_PER_TET, K_loc)]
            rhsf = [pkg.make_form(gc.IDEN, gc.P0, 1, gc.IDEN, variables[0][0], 1, 2, gc.T_NULL, gc.L_CONST)]
        else:
            _, forms, rhsf, prob = problems.c5_stokes(pkg, M, co, te)
        rp_o, ci_o, v_o, r_o, st = M.assemble(prob, co, te, dm, rank=rank)
        assert np.array_equal(da.rowptr.cpu().numpy(), rp_o) and np.array_equal(da.colind.cpu().numpy(), ci_o), "pattern not bit-exact"
        for rep in range(2):
            assert da.assemble(forms, rhsf) == 0
            torch.cuda.synchronize()
            if len(da.fields) > 1:   # vector / mixed spaces must stay on the block path under the per-rank numbering
                assert ctx.last_times()["gather_kernel"] == "k_rows_cl", "the block path did not run"
            val = da.val[:da.plan.nnz_own].cpu().numpy()
            rhs = da.rhs[:da.plan.n_own].cpu().numpy()
            rowmax = np.maximum.reduceat(np.abs(v_o), rp_o[:-1])
            rel = np.abs(val - v_o) / np.repeat(rowmax, np.diff(rp_o))
            err = rel.max()
            if not err <= 1e-12:   # diagnostic: which rows are wrong
                rows = np.repeat(np.arange(rp_o.size - 1), np.diff(rp_o))
                bad = ~(rel <= 1e-12)
                print("rank %d: %d bad entries (%d NaN) in %d rows; first rows %s; gather %s" % (
                    rank, bad.sum(), np.isnan(val).sum(), np.unique(rows[bad]).size, np.unique(rows[bad])[:10], ctx.last_times()["gather_kernel"]), flush=True)
            assert err <= 1e-12, err
            assert np.abs(rhs - r_o).max() <= 1e-12 * np.abs(r_o).max()
            if rep == 0:
                first = val.copy()
            else:
                assert np.array_equal(first, val), "multi-GPU assembly is not bit-reproducible"
        dist.barrier()
        ctx.close()
        dist.destroy_process_group()
    except Exception:
        errq.put("rank %d:\n%s" % (rank, traceback.format_exc()))


def _run(world, dims, variables):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    errq = ctx.Queue()
    port = 29700 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, dims, variables, errq)) for r in range(world)]
    for p in procs:
        p.start()
    # a rank that fails leaves its peers waiting in a collective: bound the wait and stop everybody as soon as one rank reported
    import time
    t0 = time.time()
    errs = []
    while any(p.is_alive() for p in procs) and time.time() - t0 < 300:
        while not errq.empty():
            errs.append(errq.get())
        if errs:
            time.sleep(3)
            break
        time.sleep(0.5)
    for p in procs:
        if p.is_alive():
            p.terminate()
    for p in procs:
        p.join(timeout=30)
    while not errq.empty():
        errs.append(errq.get())
    assert not errs, "\n".join(errs)
    assert all(p.exitcode == 0 for p in procs), "a rank did not finish (hung or killed)"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpus_p2(pkg, oracle):
    _run(2, (6, 4, 3), [(gc.P2, 1)])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpus_p2_many_clusters(pkg, oracle):
    """enough elements per rank for several clusters of the ring kernel in both phases of the phased assembly"""
    _run(2, (12, 10, 8), [(gc.P2, 1)])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpus_taylor_hood(pkg, oracle):
    _run(2, (4, 3, 2), [(gc.P2, 3), (gc.P1, 1)])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpus_p3(pkg, oracle):
    """P3 (face dofs, oriented edge pairs) on two GPUs.  The numbering / exchange plan is covered on the CPU
    (tests/test_parallel_cpu.py::test_two_ranks_p3); this GPU leg was added after the round's GPU budget was spent and has
    not run on hardware yet."""
    _run(2, (3, 2, 2), [(gc.P3, 1)])


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs 4 GPUs")
def test_four_gpus_p1(pkg, oracle):
    _run(4, (6, 5, 3), [(gc.P1, 1)])
