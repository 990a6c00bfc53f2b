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


def _build():
    src = os.path.join(HERE, "host", "shim_continuity.cpp")
    subprocess.check_call([
        "g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"),
        "-I" + os.path.join(PKG, "host"), src, "-o", EXE,
        "-L" + PKG, "-lnalu_edge_b200", "-Wl,-rpath," + PKG])


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
