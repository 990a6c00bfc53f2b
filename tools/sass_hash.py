#!/usr/bin/env python
"""Per-kernel SASS hash of a built library, and the comparison of two such
listings: which kernels kept their machine code across a source change.

  python tools/sass_hash.py hash LIB.so OUT.json
  python tools/sass_hash.py diff BEFORE.json AFTER.json

The hash covers the instruction text of every function (addresses and encodings
stripped).  A kernel whose template argument list was extended with a defaulted
parameter shows up under a new mangled name; `diff` pairs such renames when the
body hash is identical."""
import hashlib
import json
import re
import subprocess
import sys


def hash_lib(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], text=True,
                         capture_output=True, check=True).stdout
    out = {}
    for f in re.split(r"\n\s+Function : ", txt)[1:]:
        name = f.split("\n")[0].strip()
        body = []
        for line in f.split("\n")[1:]:
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", line)
            if m:
                body.append(m.group(1).strip())
        out[name] = [len(body), hashlib.md5("\n".join(body).encode()).hexdigest()]
    return out


def demangle(name):
    s = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    s = re.sub(r"\(anonymous namespace\)::|nw::", "", s)
    return re.sub(r"\(.*$", "", s)


def diff(a, b):
    same = [k for k in a if k in b and a[k] == b[k]]
    gone = [k for k in a if k not in b]
    new = [k for k in b if k not in a]
    changed = [k for k in a if k in b and a[k] != b[k]]
    renamed = []
    for k in list(gone):
        for n in list(new):
            if a[k] == b[n]:
                renamed.append((k, n))
                gone.remove(k)
                new.remove(n)
                break
    print("%d kernels before, %d after" % (len(a), len(b)))
    print("identical machine code under the same name: %d" % len(same))
    print("identical machine code under a new name:    %d" % len(renamed))
    for k, n in renamed:
        print("    %s  ->  %s   (%d instructions)" % (demangle(k), demangle(n), a[k][0]))
    print("changed: %d" % len(changed))
    for k in changed:
        print("    %s   %d -> %d instructions" % (demangle(k), a[k][0], b[k][0]))
    print("removed: %d" % len(gone))
    for k in gone:
        print("    %s" % demangle(k))
    print("new: %d" % len(new))
    for n in sorted(new, key=demangle):
        print("    %s   (%d instructions)" % (demangle(n), b[n][0]))
    return 0 if not changed and not gone else 1


if __name__ == "__main__":
    if sys.argv[1] == "hash":
        h = hash_lib(sys.argv[2])
        json.dump(h, open(sys.argv[3], "w"), indent=0, sort_keys=True)
        print(len(h), "functions")
    else:
        sys.exit(diff(json.load(open(sys.argv[2])), json.load(open(sys.argv[3]))))
