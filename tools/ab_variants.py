#!/usr/bin/env python
"""Build compile-time variants of the product library side by side and print the one-call A/B command.

    python tools/ab_variants.py                      # builds variants/lib_<name>.so for the candidates below
    gpurun -- '<the printed command>'                # tools/time_stages.py once per library (JSMPEG_B200_LIB)

The candidates are compile-time tuning constants (CTA sizes, occupancy bounds); round 1's macro-gated code
paths were measured and deleted (profiles/r2_variants.md).  The in-tree library is left alone.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jsmpeg_b200 import build  # noqa: E402

CANDIDATES = {
    "default": "",
    "groups4": "-DJSMPEG_EXPAND_GROUPS=4",
    "groups8": "-DJSMPEG_EXPAND_GROUPS=8",
}


def main():
    out_dir = os.path.join(ROOT, "variants")
    os.makedirs(out_dir, exist_ok=True)
    keep = build.OUT + ".keep"
    if os.path.exists(build.OUT):
        shutil.copy2(build.OUT, keep)
    cmds = []
    try:
        for name, flags in CANDIDATES.items():
            os.environ["JSMPEG_B200_NVCC_FLAGS"] = flags
            build.build(force=True)
            dst = os.path.join(out_dir, f"lib_{name}.so")
            shutil.copy2(build.OUT, dst)
            cmds.append(f"echo {name}; JSMPEG_B200_LIB=$PWD/variants/lib_{name}.so timeout 100 python tools/time_stages.py 64 60 2 2>&1 | tail -1")
    finally:
        os.environ.pop("JSMPEG_B200_NVCC_FLAGS", None)
        if os.path.exists(keep):
            os.replace(keep, build.OUT)
    print("; ".join(cmds))
    print("\n# parity first, per variant:  JSMPEG_B200_LIB=$PWD/variants/lib_<name>.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q",
          file=sys.stderr)


if __name__ == "__main__":
    main()
