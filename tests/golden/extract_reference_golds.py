#!/usr/bin/env python
"""Extract the golden vectors that pin the edge-assembly path from the
reference's own unit tests and write them to tests/golden/reference_golds.json.

Run in the build container (needs /root/reference, which does not exist on the
GPU box):  python tests/golden/extract_reference_golds.py

Only numbers are extracted (C++ brace initialisers), no code.  Sources:
  unit_tests/edge_kernels/UnitTestMomentumAdvDiffEdge.C:19-227
  unit_tests/edge_kernels/UnitTestContinuityAdvEdge.C:18-104
  unit_tests/edge_kernels/UnitTestScalarAdvDiffEdge.C:24-143
  unit_tests/ngp_algorithms/UnitTestNodalGradAlg.C:52-54, 107-110
  unit_tests/ngp_algorithms/UnitTestMdotAlg.C:66-77   (every edge == 2.5)
  unit_tests/UnitTestPecletFunction.C:36-100  (classic / tanh known answers)
"""
import json
import os
import re
import sys

REF = os.environ.get("NALU_REFERENCE", "/root/reference")


def strip_comments(s):
    s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
    return re.sub(r"//[^\n]*", "", s)


def brace_block(src, start):
    """return text of the balanced {...} starting at the first '{' >= start"""
    i = src.index("{", start)
    depth = 0
    for j in range(i, len(src)):
        if src[j] == "{":
            depth += 1
        elif src[j] == "}":
            depth -= 1
            if depth == 0:
                return src[i:j + 1]
    raise ValueError("unbalanced")


def numbers(txt):
    return [float(x) for x in re.findall(
        r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?", txt)]


def array_after(src, name):
    m = re.search(r"\b" + re.escape(name) + r"[^;{\w]*=", src)
    if not m:
        raise KeyError(name)
    return brace_block(src, m.end() - 1)


def flat(src, name):
    return numbers(array_after(src, name))


def matrix(src, name, n):
    blk = array_after(src, name)
    inner = blk[1:-1]
    rows = []
    pos = 0
    while True:
        k = inner.find("{", pos)
        if k < 0:
            break
        rb = brace_block(inner, k)
        rows.append(numbers(rb))
        pos = k + len(rb)
    assert len(rows) == n and all(len(r) == n for r in rows), (name, len(rows))
    return rows


def main():
    out = {}
    f = strip_comments(open(os.path.join(
        REF, "unit_tests/edge_kernels/UnitTestMomentumAdvDiffEdge.C")).read())
    out["momentum_adv_diff"] = {
        "rhs": flat(f, "rhs[24]"), "lhs": matrix(f, "lhs[24][24]", 24)}

    f = strip_comments(open(os.path.join(
        REF, "unit_tests/edge_kernels/UnitTestContinuityAdvEdge.C")).read())
    out["continuity_adv"] = {
        "rhs": flat(f, "rhs[8]"), "lhs": matrix(f, "lhs[8][8]", 8)}

    f = strip_comments(open(os.path.join(
        REF, "unit_tests/edge_kernels/UnitTestScalarAdvDiffEdge.C")).read())
    sc = {}
    for tag in ("serial", "P0", "P1"):
        sc[tag] = {
            "rowOffsets": [int(x) for x in flat(f, "rowOffsets_" + tag)],
            "cols": [int(x) for x in flat(f, "cols_" + tag)],
            "vals": flat(f, "vals_" + tag),
            "rhs": flat(f, "rhs_" + tag),
        }
    out["scalar_adv_diff"] = sc
    # FixPressureAtNode / applyDirichletBCs variants of the same case
    # (UnitTestScalarAdvDiffEdge.C:39-76, 108-125, 236-380)
    for kind in ("fixed", "dirichlet"):
        sc[kind + "_serial"] = {
            "vals": flat(f, kind + "_vals_serial"),
            "rhs": flat(f, kind + "_rhs_serial"),
        }

    # WallDistEdgeSolverAlg: pure 2x2 Laplacian, no rhs
    f = strip_comments(open(os.path.join(
        REF, "unit_tests/edge_kernels/UnitTestWallDistEdgeSolver.C")).read())
    out["wall_dist_edge"] = {"lhs": matrix(f, "lhs[8][8]", 8)}

    # node kernels through the same CoeffApplier boundary
    f = strip_comments(open(os.path.join(
        REF, "unit_tests/node_kernels/UnitTestScalarMassBDFNodeKernel.C")).read())
    out["scalar_mass_bdf_node"] = {
        "rhs": flat(f, "rhs[8]"), "lhs": matrix(f, "lhs[8][8]", 8)}
    f = strip_comments(open(os.path.join(
        REF, "unit_tests/node_kernels/UnitTestMomentumMassBDFNodeKernel.C")).read())
    out["momentum_mass_bdf_node"] = {"rhs": flat(f, "rhs[24]"), "lhs_diag": 1.25}
    # UnitTestContinuityMassBDFNodeKernel.C:45: expect_all_near(rhs, -12.5)
    out["continuity_mass_bdf_node"] = {"rhs_all": -12.5}

    f = strip_comments(open(os.path.join(
        REF, "unit_tests/ngp_algorithms/UnitTestNodalGradAlg.C")).read())
    i0 = f.index("NGP_nodal_grad_edge)")
    i1 = f.index("NGP_nodal_grad_edge_vec)")
    out["nodal_grad_scalar"] = flat(f[i0:i1], "expectedValues")
    out["nodal_grad_vector_diag"] = flat(f[i1:], "expectedValues")
    out["mdot_edge_value"] = 2.5

    # PecletFunction known answers (UnitTestPecletFunction.C:36-100; tolerance
    # 1e-6 there).  The inputs are C++ expressions (std::sqrt(5.0), c1 + 10.0 *
    # c2): the braces are extracted as text and evaluated with the test's own
    # constants.
    import math
    f = strip_comments(open(os.path.join(
        REF, "unit_tests/UnitTestPecletFunction.C")).read())

    def case(test_name):
        i0 = f.index(test_name + ")")
        body = f[i0:f.index("TEST(", i0) if "TEST(" in f[i0:] else len(f)]
        env = {"sqrt": math.sqrt}
        for nm in ("A", "hybridFactor", "c1", "c2"):
            m = re.search(r"\b" + nm + r"\s*=\s*([-+0-9.eE]+)\s*;", body)
            if m:
                env[nm] = float(m.group(1))

        def ev(name):
            blk = array_after(body, name)[1:-1].replace("std::sqrt", "sqrt")
            return [float(eval(x, {"__builtins__": {}}, env))
                    for x in blk.split(",")]
        d = {"peclet_numbers": ev("pecletNumbers"),
             "peclet_factors": ev("pecletFactors"), "tolerance": 1.0e-6}
        d.update({k: v for k, v in env.items() if k != "sqrt"})
        return d
    out["peclet_function"] = {
        "classic": case("NGP_classic_double"),
        "tanh": case("NGP_tanh_double"),
        "tanh_simd": case("NGP_tanh_simd"),
    }

    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                       "reference_golds.json")
    with open(dst, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", dst, {k: (len(v) if hasattr(v, "__len__") else v)
                         for k, v in out.items()})


if __name__ == "__main__":
    sys.exit(main())
