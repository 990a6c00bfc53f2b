"""The reference-named C++ host classes (nalu-wind_b200/host/NaluEdgeB200.h)
compile against the C ABI and behave like the reference's: a C++ program written
like UnitTestContinuityAdvEdge.C reproduces the reference's golden matrix."""
import json
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, "nalu-wind_b200")
EXE = os.path.join(HERE, "host", "shim_continuity")


def _build(name="shim_continuity"):
    src = os.path.join(HERE, "host", name + ".cpp")
    exe = os.path.join(HERE, "host", name)
    subprocess.check_call([
        "g++", "-std=c++17", "-O1", "-Wall", "-Wno-comment",
        "-I" + os.path.join(ROOT, "include"),
        "-I" + os.path.join(PKG, "host"), src, "-o", exe,
        "-L" + PKG, "-lnalu_edge_b200", "-Wl,-rpath," + PKG])
    return exe


def test_shim_compiles_and_refuses_without_device():
    _build()
    out = subprocess.run([EXE, "-1"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "no CPU fallback" in out.stdout


@pytest.mark.gpu
def test_shim_reproduces_continuity_gold():
    _build()
    out = subprocess.run([EXE, "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    G = json.load(open(os.path.join(HERE, "golden", "reference_golds.json")))
    rhs = None
    lhs = np.zeros((8, 8))
    for line in out.stdout.splitlines():
        t = line.split()
        if t[0] == "rhs":
            rhs = np.array([float(x) for x in t[1:]])
        elif t[0] == "lhs":
            lhs[int(t[1]), int(t[2])] = float(t[3])
    assert np.max(np.abs(rhs - np.array(G["continuity_adv"]["rhs"]))) <= 1e-12
    assert np.max(np.abs(lhs - np.array(G["continuity_adv"]["lhs"]))) <= 1e-12


def test_shim_node_program_compiles():
    _build("shim_nodes")


@pytest.mark.gpu
def test_shim_reproduces_node_walldist_geometry_golds():
    """GeometryAlgDriver, AssembleNGPNodeSolverAlgorithm::add_kernel<
    ScalarMassBDFNodeKernel / WallDistNodeKernel>, WallDistEdgeSolverAlg through
    the reference-named C++ classes, against the reference's golds"""
    exe = _build("shim_nodes")
    out = subprocess.run([exe, "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    G = json.load(open(os.path.join(HERE, "golden", "reference_golds.json")))
    mats = {"mass": np.zeros((8, 8)), "wdist": np.zeros((8, 8))}
    rhs = {}
    dnv = area = None
    for line in out.stdout.splitlines():
        t = line.split()
        if t[0] == "dnv":
            dnv = np.array([float(x) for x in t[1:]])
        elif t[0] == "area":
            area = np.array([float(x) for x in t[1:]]).reshape(12, 3)
        elif t[0].endswith("_rhs"):
            rhs[t[0][:-4]] = np.array([float(x) for x in t[1:]])
        elif t[0].endswith("_lhs"):
            mats[t[0][:-4]][int(t[1]), int(t[2])] = float(t[3])
    assert np.max(np.abs(dnv - 0.125)) <= 1e-16          # UnitTestGeometryAlg.C:64
    assert np.max(np.abs(np.sum(area * area, axis=1) - 0.0625)) <= 1e-16  # :87-97
    g = G["scalar_mass_bdf_node"]
    assert np.max(np.abs(mats["mass"] - np.array(g["lhs"]))) <= 1e-12
    assert np.max(np.abs(rhs["mass"] - np.array(g["rhs"]))) <= 1e-12
    assert np.max(np.abs(mats["wdist"] - np.array(G["wall_dist_edge"]["lhs"]))) <= 1e-12
    assert np.max(np.abs(rhs["wdist"] - 0.125)) <= 1e-16  # WallDistNodeKernel: rhs += V
