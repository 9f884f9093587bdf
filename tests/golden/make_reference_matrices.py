#!/usr/bin/env python
"""Transcribes the reference's FIXED regression inputs into tests/golden/reference_matrices.json.

Run in the build container (needs /root/reference; the GPU box does not have it — the JSON is the
committed fixture).  Sources (PetrKryslUCSD/Sparspak.jl, test/):
  * test_structunsymm.jl:13-21   simpletest1: I(4) with A[2,1] = -0.1          (used to fail in issymmetric)
  * test_structunsymm.jl:26-41   simpletest2: 4x4, used to fail in SpkLUFactor.jl:245
  * test_structunsymm.jl:46-60   simpletest3: 4x4, used to fail in SpkSparseBase.jl:390
  * test_small.jl:35-59          m_simpletest1: the simpletest2 matrix entered with inaij!, rhs e1
  * test_small.jl:70-104         symmetric structure, unsymmetric values
  * test_small.jl:115-149        unsymmetric structure
  * test_sparse_method.jl:389-424 the 31x31 "random" matrix (sparse(I, J, V, 31, 31)), rhs 1:31
The expected results in the reference are solves against dense `\\` to 1e-6; the fixtures therefore hold the
inputs, and tests/ compares oracle, dense solve and the CUDA path on them."""
import json
import os
import re

REF = "/root/reference/test"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_matrices.json")


def dense_block(text, start_pat):
    """Parse a Julia dense literal `A=[ a b c; ... ]` following start_pat."""
    m = re.search(start_pat + r"\s*A\s*=\s*\[(.*?)\]", text, re.S)
    rows = [r.split() for r in m.group(1).replace("\n", " ").split(";") if r.strip()]
    return [[float(v) for v in r] for r in rows]


def triplets(A):
    I, J, V = [], [], []
    for j in range(len(A[0])):
        for i in range(len(A)):
            if A[i][j] != 0.0:
                I.append(i + 1); J.append(j + 1); V.append(A[i][j])
    return I, J, V


def main():
    su = open(os.path.join(REF, "test_structunsymm.jl")).read()
    sm = open(os.path.join(REF, "test_small.jl")).read()
    sp = open(os.path.join(REF, "test_sparse_method.jl")).read()
    out = {"_provenance": __doc__}

    def put(name, source, n, I, J, V, rhs):
        out[name] = {"source": source, "n": n, "I": I, "J": J, "V": V, "rhs": rhs}

    put("structunsymm_simpletest1", "test/test_structunsymm.jl:13-21", 4, [1, 2, 2, 3, 4], [1, 1, 2, 3, 4],
        [1.0, -0.1, 1.0, 1.0, 1.0], [1.0, 0.0, 0.0, 0.0])
    A2 = dense_block(su, r"function simpletest2\(\)")
    put("structunsymm_simpletest2", "test/test_structunsymm.jl:26-41", 4, *triplets(A2), [1.0, 0.0, 0.0, 0.0])
    A3 = dense_block(su, r"function simpletest3\(\)")
    put("structunsymm_simpletest3", "test/test_structunsymm.jl:46-60", 4, *triplets(A3), [1.0, 0.0, 0.0, 0.0])
    # test_small.jl m_simpletest1: inaij! calls
    blk = sm[sm.index("module m_simpletest1"): sm.index("module m_inconsistency_data_structure_symm")]
    ent = re.findall(r"inaij!\(p,\s*(\d+),\s*(\d+),\s*([-0-9.eE]+)\)", blk)
    put("small_simpletest1", "test/test_small.jl:35-59", 4, [int(e[0]) for e in ent], [int(e[1]) for e in ent],
        [float(e[2]) for e in ent], [1.0, 0.0, 0.0, 0.0])
    blk = sm[sm.index("module m_inconsistency_data_structure_symm"): sm.index("module m_inconsistency_data_structure_unsymm")]
    put("small_symm_structure", "test/test_small.jl:70-104", 4, *triplets(dense_block(blk, r"n = 4")), [1.0, 0.0, 0.0, 0.0])
    blk = sm[sm.index("module m_inconsistency_data_structure_unsymm"):]
    put("small_unsymm_structure", "test/test_small.jl:115-149", 4, *triplets(dense_block(blk, r"n = 4")), [1.0, 0.0, 0.0, 0.0])
    # 31x31
    m = re.search(r"spm = sparse\(\[(.*?)\],\s*\[(.*?)\],\s*T\[(.*?)\],\s*31,\s*31\)", sp, re.S)
    I = [int(v) for v in m.group(1).replace("\n", " ").split(",")]
    J = [int(v) for v in m.group(2).replace("\n", " ").split(",")]
    V = [float(v) for v in m.group(3).replace("\n", " ").split(",")]
    assert len(I) == len(J) == len(V)
    put("sparse_method_31x31", "test/test_sparse_method.jl:389-424", 31, I, J, V, [float(i) for i in range(1, 32)])
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT, {k: len(v["I"]) for k, v in out.items() if k != "_provenance"})


if __name__ == "__main__":
    main()
