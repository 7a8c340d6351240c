#!/usr/bin/env python
"""md5 of the SASS instruction stream of the product kernels (per .cu file), addresses and encodings
stripped.  Used to show that a source refactoring left the machine code of a GPU-verified kernel
untouched (profiles/README.md lists the hashes of the build the round-end numbers were taken with).

    python tools/sass_hash.py [extra nvcc flags...]
"""
import hashlib
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "jsmpeg_b200", "csrc")


def sass_hash(cu, flags=()):
    with tempfile.TemporaryDirectory() as tmp:
        cubin = os.path.join(tmp, "k.cubin")
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", *flags,
                               "-cubin", "-o", cubin, os.path.join(CSRC, cu)])
        sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True, check=True).stdout
    lines = [l for l in sass.splitlines() if re.match(r"^\s+/\*[0-9a-f]{4,5}\*/", l)]
    return hashlib.md5(("\n".join(lines) + "\n").encode()).hexdigest(), len(lines)


if __name__ == "__main__":
    for cu in ("parse.cu", "recon.cu", "scan.cu", "tsdemux.cu"):
        h, n = sass_hash(cu, sys.argv[1:])
        print(f"{cu:12s} {h}  ({n} instructions)")
