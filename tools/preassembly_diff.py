#!/usr/bin/env python
"""Reader / differ for nalu-wind's pre-assembly dumps (solver option
write_preassembly_matrix_files; writers src/HypreLinearSystem.C:1517-1568,
1625-1661 and src/HypreUVWLinearSystem.C:135-166 of the reference, and
nw_linsys_write_preassembly_files of this library):

    <eq>.IJM.<n>.mat.<rank%05d>.preassem.{i,j,v,meta}
    <eq>[d].IJV.<n>.rhs.<rank%05d>.preassem.{i,v,meta}

  python tools/preassembly_diff.py DIR_A DIR_B [--rtol 1e-12] [--eq NAME]

compares every dump present in both directories: integer files (.i .j .meta)
must be identical (the integer width of each side is taken from its .meta
size, so a 32-bit and a bigint build can be compared); values within --rtol of
max(|a|, |b|, row scale), row scale = largest |entry| of the row (the reference
sums with atomics, so the last bits differ from run to run).  If the two sides
list the same entries in a different order the comparison is done on the
(row, col)-sorted, duplicate-summed form and the re-ordering is reported.
Exit status 0: all common dumps agree; 1: a difference; 2: nothing to compare.
numpy only."""
import argparse
import glob
import os
import sys

import numpy as np


def _ints(path, n_meta_words=None):
    """read an integer file; width from the companion .meta (6 words for a
    matrix, 3 for a vector)"""
    base = path[: path.rindex(".")]
    words = 6 if ".mat." in base else 3
    msize = os.path.getsize(base + ".meta")
    if msize not in (4 * words, 8 * words):
        raise ValueError("%s.meta: %d bytes is neither %d 32-bit nor 64-bit words"
                         % (base, msize, words))
    dt = np.int32 if msize == 4 * words else np.int64
    return np.fromfile(path, dtype=dt).astype(np.int64)


def read_matrix(base):
    """base = path up to and including '.preassem'; returns dict(i, j, v, meta)"""
    d = dict(i=_ints(base + ".i"), j=_ints(base + ".j"),
             v=np.fromfile(base + ".v", dtype=np.float64), meta=_ints(base + ".meta"))
    if not (len(d["i"]) == len(d["j"]) == len(d["v"])):
        raise ValueError(base + ": .i/.j/.v lengths differ")
    if len(d["meta"]) != 6 or d["meta"][5] != len(d["v"]) or \
            d["meta"][3] + d["meta"][4] != d["meta"][5]:
        raise ValueError(base + ": meta does not describe the arrays")
    return d


def read_vector(base):
    d = dict(i=_ints(base + ".i"), v=np.fromfile(base + ".v", dtype=np.float64),
             meta=_ints(base + ".meta"))
    if len(d["meta"]) != 3 or d["meta"][0] + d["meta"][1] != d["meta"][2]:
        raise ValueError(base + ": meta does not describe the arrays")
    return d


def _canonical(i, j, v):
    """(row, col)-sorted, duplicates summed"""
    order = np.lexsort((j, i))
    i, j, v = i[order], j[order], v[order]
    new = np.ones(len(i), dtype=bool)
    new[1:] = (i[1:] != i[:-1]) | (j[1:] != j[:-1])
    idx = np.cumsum(new) - 1
    out = np.zeros(int(new.sum()))
    np.add.at(out, idx, v)
    return i[new], j[new], out


def _value_report(rows, a, b, rtol):
    scale = np.maximum(np.abs(a), np.abs(b))
    if len(rows):
        r0 = rows - rows.min()
        rowmax = np.zeros(int(r0.max()) + 1)
        np.maximum.at(rowmax, r0, scale)
        scale = np.maximum(scale, rowmax[r0])
    scale = np.maximum(scale, 1e-300)
    err = np.abs(a - b) / scale
    k = int(np.argmax(err)) if len(err) else -1
    worst = float(err[k]) if k >= 0 else 0.0
    return worst <= rtol, worst, k, int(np.count_nonzero(a != b))


def diff_matrix(a, b, rtol):
    msgs, ok = [], True
    if not np.array_equal(a["meta"], b["meta"]):
        ok = False
        msgs.append("meta differs: %s vs %s" % (a["meta"].tolist(), b["meta"].tolist()))
    ia, ja, va, ib, jb, vb = a["i"], a["j"], a["v"], b["i"], b["j"], b["v"]
    if not (np.array_equal(ia, ib) and np.array_equal(ja, jb)):
        ia, ja, va = _canonical(ia, ja, va)
        ib, jb, vb = _canonical(ib, jb, vb)
        if not (np.array_equal(ia, ib) and np.array_equal(ja, jb)):
            return False, msgs + ["different sparsity: %d vs %d distinct entries"
                                  % (len(ia), len(ib))]
        msgs.append("same entries in a different order (compared sorted)")
    good, worst, k, nbits = _value_report(ia, va, vb, rtol)
    msgs.append("%d entries, %d not bit-identical, worst scaled difference %.3e%s"
                % (len(va), nbits, worst,
                   "" if k < 0 else " at (row %d, col %d)" % (ia[k], ja[k])))
    return ok and good, msgs


def diff_vector(a, b, rtol):
    msgs, ok = [], True
    if not np.array_equal(a["meta"], b["meta"]):
        ok = False
        msgs.append("meta differs: %s vs %s" % (a["meta"].tolist(), b["meta"].tolist()))
    if not np.array_equal(a["i"], b["i"]):
        return False, msgs + ["row ids differ"]
    n = min(len(a["v"]), len(b["v"]))
    if len(a["v"]) != len(b["v"]):
        ok = False
        msgs.append("lengths differ: %d vs %d" % (len(a["v"]), len(b["v"])))
    va, vb = a["v"][:n], b["v"][:n]
    scale = max(float(np.max(np.abs(va))) if n else 0.0,
                float(np.max(np.abs(vb))) if n else 0.0, 1e-300)
    err = np.abs(va - vb) / scale
    worst = float(err.max()) if n else 0.0
    msgs.append("%d rows, %d not bit-identical, worst difference %.3e of the largest entry"
                % (n, int(np.count_nonzero(va != vb)), worst))
    return ok and worst <= rtol, msgs


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("dir_a")
    ap.add_argument("dir_b")
    ap.add_argument("--rtol", type=float, default=1e-12)
    ap.add_argument("--eq", default="*", help="equation system name (glob)")
    args = ap.parse_args(argv)
    found, bad = 0, 0
    for kind, reader, differ in (("mat", read_matrix, diff_matrix),
                                 ("rhs", read_vector, diff_vector)):
        pat = "%s.IJ%s.*.%s.*.preassem.meta" % (args.eq, "M" if kind == "mat" else "V", kind)
        for meta in sorted(glob.glob(os.path.join(args.dir_a, pat))):
            base = os.path.basename(meta)[: -len(".meta")]
            other = os.path.join(args.dir_b, base)
            if not os.path.exists(other + ".meta"):
                continue
            found += 1
            try:
                ok, msgs = differ(reader(os.path.join(args.dir_a, base)), reader(other),
                                  args.rtol)
            except (ValueError, OSError) as e:
                ok, msgs = False, [str(e)]
            bad += 0 if ok else 1
            print("%s %s" % ("ok  " if ok else "DIFF", base))
            for m in msgs:
                print("     " + m)
    if not found:
        print("no dump present in both directories")
        return 2
    print("%d dumps compared, %d differ" % (found, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
