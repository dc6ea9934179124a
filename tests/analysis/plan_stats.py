#!/usr/bin/env python3
"""CPU emulation of the slice builder of k_rows_cl (inmost-fem_b200/csrc/afb_rows.cu: k_morton, k_row_key, slices of 32 rows per
(cluster, length bucket), visit-steps = sum over classes of the slice maximum) to put numbers on plan-level levers before they
are built.  TEST-SIDE ANALYSIS TOOL (lives under tests/ because it uses the oracle's mesh / numbering helpers); not part of the product.

  python tests/analysis/plan_stats.py [--n 40] [--chunk 512]

Reports the fraction of real visits for
  * the current plan (10 visit classes for P2 = local row index),
  * merged classes (vertex rows / edge rows: the per-visit permutation of the six off-diagonal G entries, profiles/r01e_*.md lever 4),
  * merged classes + slices cut per cluster only (no length buckets).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

BUCKETS = [20, 28, 36, 48, 66, 96, 128, 192, 256]


def spread3(v):
    v = v & 0x3ff
    v = (v | (v << 16)) & 0x030000ff
    v = (v | (v << 8)) & 0x0300f00f
    v = (v | (v << 4)) & 0x030c30c3
    v = (v | (v << 2)) & 0x09249249
    return v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=40)
    ap.add_argument("--chunk", type=int, default=512)
    args = ap.parse_args()
    O, M = entry.load_oracle()
    co, te, _ = M.cube_mesh(args.n, args.n, args.n)
    dm = M.DofMap(te, [(O.P2, 1)], nnode=co.shape[0])
    ntet, nloc = dm.elem2dof.shape
    # Morton order of the centroids (k_morton)
    c = co[te].mean(axis=1)
    lo, ext = co.min(axis=0), co.max(axis=0) - co.min(axis=0)
    q = np.clip(((c - lo) / ext * 1024).astype(np.int64), 0, 1023)
    code = spread3(q[:, 0]) | (spread3(q[:, 1]) << 1) | (spread3(q[:, 2]) << 2)
    new2old = np.argsort(code, kind="stable")
    old2new = np.empty(ntet, dtype=np.int64)
    old2new[new2old] = np.arange(ntet)
    # adjacency: visits (row, element, i)
    rows = dm.elem2dof.ravel()
    el = np.repeat(np.arange(ntet), nloc)
    li = np.tile(np.arange(nloc), ntet)
    nrows = dm.nrows
    # row length = distinct columns
    rr = np.repeat(dm.elem2dof, nloc, axis=1).ravel()
    cc = np.tile(dm.elem2dof, (1, nloc)).ravel()
    key = np.unique(rr.astype(np.int64) * nrows + cc)
    rowlen = np.bincount(key // nrows, minlength=nrows)
    bucket = np.searchsorted(np.array(BUCKETS), rowlen, side="left").clip(max=len(BUCKETS) - 1)
    cl = np.full(nrows, np.iinfo(np.int64).max)
    np.minimum.at(cl, rows, old2new[el])
    cl //= args.chunk
    cnt = np.zeros((nrows, nloc), dtype=np.int64)
    np.add.at(cnt, (rows, li), 1)
    deg = cnt.sum(axis=1)
    nadj = int(deg.sum())

    def steps(class_counts, use_bucket=True):
        pack = np.zeros(nrows, dtype=np.int64)
        bits = max(1, min(6, 30 // class_counts.shape[1]))
        for k in range(class_counts.shape[1]):
            pack = (pack << bits) | np.minimum(class_counts[:, k], (1 << bits) - 1)
        grp = cl * 16 + (bucket if use_bucket else 0)
        order = np.lexsort((pack, np.minimum(deg, 63), grp))
        g = grp[order]
        start = np.r_[True, g[1:] != g[:-1]]
        gstart = np.maximum.accumulate(np.where(start, np.arange(nrows), 0))
        pos = np.arange(nrows) - gstart
        slice_id = np.cumsum((pos % 32) == 0) - 1
        ns = slice_id[-1] + 1
        total = 0
        for k in range(class_counts.shape[1]):
            m = np.zeros(ns, dtype=np.int64)
            np.maximum.at(m, slice_id, class_counts[order, k])
            total += int(m.sum())
        return ns, total

    print("cube %d^3: %d tets, %d rows, %d visits, chunk %d" % (args.n, ntet, nrows, nadj, args.chunk))
    for name, cc_, ub in (("current plan (10 classes)", cnt, True),
                          ("vertex / edge classes merged", np.stack([cnt[:, :4].sum(1), cnt[:, 4:].sum(1)], 1), True),
                          ("merged classes, slices per cluster only", np.stack([cnt[:, :4].sum(1), cnt[:, 4:].sum(1)], 1), False)):
        ns, tot = steps(cc_, ub)
        print("  %-42s %7d slices (%.1f%% lanes filled), %8d visit-steps, %.1f%% real visits" % (name, ns, 100.0 * nrows / (32 * ns), tot, 100.0 * nadj / (32 * tot)))


if __name__ == "__main__":
    main()
