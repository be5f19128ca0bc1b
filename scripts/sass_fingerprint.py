#!/usr/bin/env python
"""Per-kernel SASS fingerprints of the built library: `python scripts/sass_fingerprint.py out.json`.
Used to show that an edit left the machine code of already GPU-validated kernels untouched (addresses and the library's own
symbol hashes are normalised away; everything else -- opcodes, registers, immediates, predicates -- must match)."""
import hashlib
import json
import re
import subprocess
import sys

import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from orb_slam3_detailed_comments_b200 import _native as N  # noqa: E402


def fingerprints(lib=N.LIB_PATH):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, name, buf = {}, None, []
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                out[name] = buf
            name, buf = m.group(1), []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?)\s*/\* 0x[0-9a-f]+ \*/", line)
        if m and name:
            buf.append(re.sub(r"0x[0-9a-f]{5,}", "ADDR", m.group(1)))       # branch targets / constant-bank addresses of the image
    if name:
        out[name] = buf
    norm = lambda n: re.sub(r"_GLOBAL__N__[0-9a-f]+_", "_GLOBAL__N__", n)   # anonymous-namespace hash depends on the TU
    return {norm(k): (len(v), hashlib.sha256("\n".join(v).encode()).hexdigest()[:16]) for k, v in out.items()}


if __name__ == "__main__":
    fp = fingerprints()
    json.dump(fp, open(sys.argv[1], "w"), indent=0, sort_keys=True)
    print(len(fp), "kernels")
