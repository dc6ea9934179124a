#!/usr/bin/env python3
"""Regenerate the committed golden fixtures of tests/golden/ (run in the build container only).

Two kinds of fixture, both taken from the reference (/root/reference, read-only):

 (1) known-answer tables the reference's own unit tests hold for this path, extracted verbatim from
     the test sources:
       tests/fem/operations/int_tet_test.cpp:230-244   GRAD(P3) x GRAD(P1^3), polynomial GENERAL tensor, /720
       tests/fem/operations/int_tet_test.cpp:334-347   GRAD(P1^3)^2, identity tensor, /1440
       tests/fem/operations/int_tet_test.cpp:455       rhs trick IDEN(P0) x IDEN(P2^3): {-1 x4, 4 x6} x3
       tests/fem/spaces/predefined_spaces_test.cpp:62-382   U tables of IDEN/GRAD on P0..P3, f = 2 tets
     -> reference_tests.json
 (1b) tests/fem/operations/int_face_test.cpp:78-89   fem3Dface GRAD(P2) x IDEN(P1^3), face 1, normal-weighted tensor
 (2) outputs of the reference itself (oracle/_ref/libanifem_ref.so, built by oracle/Makefile from the
     unmodified sources) on seeded inputs for every operator/space/tensor family of SURVEY.md section 8a
     -> ref_outputs.npz  (inputs are regenerated from the seed by tests/golden_cases.py)

    python tests/golden/make_golden.py
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def numbers(text):
    return [float(x) for x in re.findall(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?", text)]


def brace_block(src, start):
    """text of the {...} block beginning at the first '{' at or after `start`"""
    i = src.index("{", start)
    depth, j = 0, i
    while True:
        if src[j] == "{":
            depth += 1
        elif src[j] == "}":
            depth -= 1
            if depth == 0:
                return src[i + 1:j], j
        j += 1


def extract_int_tet():
    src = open(os.path.join(REF, "tests/fem/operations/int_tet_test.cpp")).read()
    out = {}
    k = src.index("At_exp_m{std::array<double, 20>")
    blk, _ = brace_block(src, k)
    blk = re.sub(r"std::array<double, 20>", "", blk)
    vals = numbers(blk)
    assert len(vals) == 12 * 20
    out["grad_p3_x_grad_p1vec_general"] = {"coef": 720.0, "table_rows_test_cols_trial": np.array(vals).reshape(12, 20).tolist(),
                                           "tet": [[0, 0, 0], [2, 1, 1], [1, 2, 1], [2, 1, 2]], "order": 5,
                                           "tensor": "D(i,j) = (i*x0 + j*x1 + (i%2)*x2) * (i*ncols + j), 9x3 GENERAL"}
    k = src.index("double A_exp_m1p[]")
    blk, _ = brace_block(src, k)
    vals = numbers(blk)
    assert len(vals) == 144
    out["grad_p1vec_sq_identity"] = {"coef": 1440.0, "table": np.array(vals).reshape(12, 12).tolist(),
                                     "tet": [[0, 0, 0], [2, 1, 1], [1, 2, 1], [2, 1, 2]], "order": 5}
    k = src.index("TEST(AniInterface, RhsEval)")
    k = src.index("Bp[30]", k)
    blk, _ = brace_block(src, k)
    vals = numbers(blk)
    assert len(vals) == 30
    out["rhs_p0_x_iden_p2vec"] = {"table": vals, "mu": 40.0, "tet": [[0, 0, 0], [2, 1, 1], [1, 2, 1], [2, 1, 2]], "order": 2}
    return out


def extract_int_face():
    """tests/fem/operations/int_face_test.cpp:78-89: GRAD(P2) x IDEN(P1^3) over face 1 of the tet (1,1,1),(2,1,1),(1,2,1),(1,1,2),
    order 3, tensor D(i,j) = sum_k (x[j%3] + (3i+k)/10) n_k with n the outward normal of the face (:29-56)"""
    src = open(os.path.join(REF, "tests/fem/operations/int_face_test.cpp")).read()
    k = src.index("A_expd = {")
    vals = numbers(brace_block(src, k)[0])
    assert len(vals) == 120
    return {"grad_p2_x_iden_p1vec_face1": {"table_cols_trial_rows_test": np.array(vals).reshape(10, 12).tolist(), "face": 1, "order": 3,
                                           "tet": [[1, 1, 1], [2, 1, 1], [1, 2, 1], [1, 1, 2]],
                                           "tensor": "D(i,j) = sum_k (x[j%3] + (3i+k)/10) n_k, 3x3 GENERAL, n = outward normal"}}


def extract_spaces():
    src = open(os.path.join(REF, "tests/fem/spaces/predefined_spaces_test.cpp")).read()
    end = src.index("#undef SETOP")
    head = src[:end]
    q, fusion = 4, 2
    k = head.index("double XYL[4*q]")
    XYL = numbers(brace_block(head, k)[0])
    k = head.index("double XYZ[3*fusion*4]")
    blk = re.sub(r"//.*", "", brace_block(head, k)[0])
    XYZ = numbers(blk)
    assert len(XYL) == 16 and len(XYZ) == 24
    out = {"XYL": XYL, "XYZ_layout": "XYZ[3*r + 3*fusion*l + k] = coordinate k of vertex l of tet r", "XYZ": XYZ, "tables": {}}
    for m in re.finditer(r"SETOP\((IDEN|GRAD), (FEM_P[0-3])\);", head):
        op, fem = m.group(1), m.group(2)
        # enclosing block: last "    {" before the match
        b0 = head.rfind("\n    {", 0, m.start())
        body = head[b0:m.start()]
        dims = re.search(r"nfa_exp = (\d+), dim_exp = ([\d\*]+)", body)
        nfa, dim = int(dims.group(1)), eval(dims.group(2))
        n1 = dim * q * nfa
        if "double Ut[]" in body:
            vals = numbers(brace_block(body, body.index("double Ut[]"))[0])
            assert len(vals) == n1, (op, fem, len(vals), n1)
            vals = vals * fusion
        elif re.search(r"double U_exp\[\]\s*=", body):
            mm = re.search(r"double U_exp\[\]\s*=", body)
            vals = numbers(brace_block(body, mm.start())[0])
            assert len(vals) == n1 * fusion, (op, fem, len(vals), n1 * fusion)
        else:  # GRAD P0: zero-filled
            vals = [0.0] * (n1 * fusion)
        out["tables"]["%s_%s" % (op, fem)] = {"nfa": nfa, "dim": dim, "U": vals}
    return out


def main():
    ref_tests = {"int_tet": extract_int_tet(), "int_face": extract_int_face(), "spaces": extract_spaces()}
    with open(os.path.join(HERE, "reference_tests.json"), "w") as f:
        json.dump(ref_tests, f)
    print("reference_tests.json:", list(ref_tests["int_tet"].keys()), list(ref_tests["spaces"]["tables"].keys()))
    # (2) reference outputs on the seeded cases
    import golden_cases
    from oracle import oracle as O
    assert O.have_ref(), "build oracle/_ref first: make -C oracle ref"
    out = {}
    for name, form, XY, D in golden_cases.cases():
        out[name] = O.fem3dtet(form, XY, D, impl="ref", mode=0)
        tmpl = O.fem3dtet(form, XY, D, impl="ref", mode=1, fuse=2)
        assert np.abs(tmpl - out[name]).max() <= 1e-13 * (1 + np.abs(out[name]).max()), name
    np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **out)
    print("ref_outputs.npz:", len(out), "cases")
    # (3) the reference's fem3Dface on the seeded surface cases
    outf = {}
    for name, form, XY, face, D in golden_cases.face_cases():
        outf[name] = O.fem3dface(form, XY, face, D, impl="ref")
    np.savez_compressed(os.path.join(HERE, "ref_face_outputs.npz"), **outf)
    print("ref_face_outputs.npz:", len(outf), "cases")


if __name__ == "__main__":
    main()
