"""Generates tests/golden/ref_assembler.npz: outputs of the reference's OWN global assembler (unmodified
anifem++/inmost_interface sources on oracle/mock_inmost, built by `make -C oracle refasm` where /root/reference exists) for
small meshes: dof index codes of every GlobEnumeration type, the AssembleTemplate pattern, and assembled matrices / load vectors.
tests/test_oracle_golden.py compares oracle/asm_oracle.py with these fixtures (and with the live build when it is present).
Run from the repo root:  python tests/golden/make_golden_asm.py"""
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import asm_oracle as M, oracle as O   # noqa: E402
import asm_cases   # noqa: E402


def main():
    out = {}
    for name, co, te, variables in asm_cases.numbering_cases(M):
        for et in O.RefAssembler.ENUM:
            R = O.RefAssembler(co, te, variables, et)
            out["num/%s/%s/codes" % (name, et)] = R.codesC
            out["num/%s/%s/nrows" % (name, et)] = np.array([R.nrows])
            assert np.array_equal(R.codesC, R.codesR)
    for name, co, te, variables, prob, kw in asm_cases.assembly_cases(M, O):
        R = O.RefAssembler(co, te, variables, "NATURAL")
        rp, ci = R.template()
        out["asm/%s/template_rowptr" % name], out["asm/%s/template_colind" % name] = rp, ci
        for mode, opts in (("plain", {}), ("templ", dict(include_template=True, ordered_insert=True))):
            st, rp2, ci2, v, r = R.assemble(prob, drop_val=kw.get("drop_val", 1e-100), **opts)
            out["asm/%s/%s/status" % (name, mode)] = np.array([st])
            out["asm/%s/%s/rowptr" % (name, mode)], out["asm/%s/%s/colind" % (name, mode)] = rp2, ci2
            out["asm/%s/%s/val" % (name, mode)], out["asm/%s/%s/rhs" % (name, mode)] = v, r
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_assembler.npz"), **out)
    print("wrote %d arrays" % len(out))


if __name__ == "__main__":
    main()
