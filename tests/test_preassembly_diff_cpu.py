"""tools/preassembly_diff.py: the reader / differ of the reference's
pre-assembly dump format (src/HypreLinearSystem.C:1517-1568, 1625-1661),
exercised on synthetic dumps written with numpy in the writer's layout (the
library's own writer is checked against this layout on the GPU by
tests/test_gpu_parity.py::test_preassembly_files_round_trip).  No GPU."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location(
    "preassembly_diff", os.path.join(HERE, "..", "tools", "preassembly_diff.py"))
pd = importlib.util.module_from_spec(spec)
spec.loader.exec_module(pd)


def _write(d, eq, rows, cols, vals, rhs, it, n_owned, counter=3, rank=0):
    os.makedirs(d, exist_ok=True)
    base = os.path.join(d, "%s.IJM.%d.mat.%05d.preassem." % (eq, counter, rank))
    rows.astype(it).tofile(base + "i")
    cols.astype(it).tofile(base + "j")
    vals.tofile(base + "v")
    nnz = len(vals)
    np.array([50, 0, 49, n_owned, nnz - n_owned, nnz], dtype=it).tofile(base + "meta")
    vb = os.path.join(d, "%s.IJV.%d.rhs.%05d.preassem." % (eq, counter, rank))
    np.arange(len(rhs)).astype(it).tofile(vb + "i")
    rhs.tofile(vb + "v")
    np.array([len(rhs) - 2, 2, len(rhs)], dtype=it).tofile(vb + "meta")


def _system(seed=1):
    rng = np.random.default_rng(seed)
    rows = np.repeat(np.arange(50), 5)
    cols = (rows + np.tile(np.arange(-2, 3), 50)) % 50
    return rows, cols, rng.standard_normal(len(rows)), rng.standard_normal(50)


def test_identical_across_integer_widths(tmp_path, capsys):
    r, c, v, b = _system()
    _write(tmp_path / "a", "ContinuityEQS", r, c, v, b, np.int32, 200)
    _write(tmp_path / "b", "ContinuityEQS", r, c, v, b, np.int64, 200)
    assert pd.main([str(tmp_path / "a"), str(tmp_path / "b")]) == 0
    out = capsys.readouterr().out
    assert "2 dumps compared, 0 differ" in out and "0 not bit-identical" in out
    m = pd.read_matrix(str(tmp_path / "b" / "ContinuityEQS.IJM.3.mat.00000.preassem"))
    assert m["meta"].tolist() == [50, 0, 49, 200, 50, 250]
    assert np.array_equal(m["i"], r) and np.array_equal(m["v"], v)


def test_rounding_passes_perturbation_fails(tmp_path, capsys):
    r, c, v, b = _system(2)
    _write(tmp_path / "a", "MomentumEQS0", r, c, v, b, np.int32, 200)
    v2 = v * (1.0 + 3e-16 * np.sign(np.sin(np.arange(len(v)))))
    _write(tmp_path / "b", "MomentumEQS0", r, c, v2, b * (1 + 2e-16), np.int32, 200)
    assert pd.main([str(tmp_path / "a"), str(tmp_path / "b")]) == 0
    v3 = v.copy()
    v3[17] += 1e-9
    _write(tmp_path / "c", "MomentumEQS0", r, c, v3, b, np.int32, 200)
    assert pd.main([str(tmp_path / "a"), str(tmp_path / "c")]) == 1
    out = capsys.readouterr().out
    assert "DIFF MomentumEQS0.IJM.3.mat.00000.preassem" in out
    assert "(row %d, col %d)" % (r[17], c[17]) in out


def test_reordered_entries_and_sparsity_change(tmp_path, capsys):
    r, c, v, b = _system(3)
    _write(tmp_path / "a", "EnthalpyEQS", r, c, v, b, np.int64, 200)
    p = np.random.default_rng(0).permutation(len(v))
    _write(tmp_path / "b", "EnthalpyEQS", r[p], c[p], v[p], b, np.int64, 200)
    assert pd.main([str(tmp_path / "a"), str(tmp_path / "b")]) == 0
    assert "different order" in capsys.readouterr().out
    c2 = c.copy()
    c2[5] = (c2[5] + 7) % 50
    _write(tmp_path / "c", "EnthalpyEQS", r, c2, v, b, np.int64, 200)
    assert pd.main([str(tmp_path / "a"), str(tmp_path / "c")]) == 1
    assert "different sparsity" in capsys.readouterr().out
    assert pd.main([str(tmp_path / "a"), str(tmp_path / "nothing")]) == 2
