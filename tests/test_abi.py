"""CPU tests (no GPU): the C-ABI library loads, exports every symbol include/anifem_b200.h declares,
its host-only entry points work, and compute entry points fail loudly without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

import golden_cases as gc
from conftest import ROOT


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "anifem_b200.h")).read()
    declared = set(re.findall(r"\b(afb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared == set(pkg.EXPORTS)
    L = pkg.lib()
    for sym in declared:
        assert getattr(L, sym) is not None


def test_op_dims(pkg):
    for op in (gc.IDEN, gc.GRAD, gc.DIV):
        for fem in (gc.P0, gc.P1, gc.P2, gc.P3):
            for vec in (1, 3):
                if op == gc.DIV and (vec != 3 or fem == gc.P0):
                    with pytest.raises(pkg.AfbError):
                        pkg.op_dims(op, fem, vec)
                else:
                    assert pkg.op_dims(op, fem, vec) == gc.op_dims(op, fem, vec)
    with pytest.raises(pkg.AfbError):
        pkg.op_dims(gc.IDEN, 21, 1)  # FEM_RT0 is out of scope


def test_quadrature_tables_match_oracle(pkg, oracle):
    for order in range(0, 21):
        p, w = pkg.tet_quadrature(order)
        po, wo = oracle.tet_quadrature(order)
        assert np.array_equal(p, po) and np.array_equal(w, wo)
    with pytest.raises(pkg.AfbError):
        pkg.tet_quadrature(21)
    for order in range(0, 21):   # triangle rules of fem3Dface
        p, w = pkg.tri_quadrature(order)
        po, wo = oracle.tri_quadrature(order)
        assert np.array_equal(p, po) and np.array_equal(w, wo)
    with pytest.raises(pkg.AfbError):
        pkg.tri_quadrature(21)


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.AfbError) as e:
        pkg.Context(0)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)


def test_new_entries_reject_a_missing_context(pkg):
    """the entries added in round 2 (MatFuncWrap scatter, communicator, halo plan / exchange, distributed assembly, kernel names)
    return -7 without touching a device when no context is given -- no CPU path behind them"""
    import ctypes
    L = pkg.lib()
    assert L.afb_assemble_elemental(None, 0, 1, None, None, 0, None, None, 1e-100, 0) == -7
    assert L.afb_comm_init(None, None, 0, 1) == -7
    assert L.afb_comm_set(None, None, 0, 1) == -7
    assert L.afb_halo_plan_set(None, 1, 0, 0, None, None, None, None, None, None, 0) == -7
    assert L.afb_halo_exchange_start(None, None, None) == -7
    assert L.afb_halo_exchange_finish(None, None, None) == -7
    assert L.afb_halo_exchange(None, None, None) == -7
    assert L.afb_assemble_distributed(None, 0, None, 0, None, None, None, 1e-100) == -7
    buf = ctypes.create_string_buffer(16)
    assert L.afb_last_kernels(None, buf, 16) == -7
    assert L.afb_comm_unique_id(None) == -7
