"""Build libgravinv_b200.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build()."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", f) for f in ("assemble.cu", "leapfrog.cu", "batched.cu", "wavelet.cu", "reginv.cu", "sink.cu", "fields.cu", "tess_fields.cu", "fused.cu", "hostrng.cu")]
HDR = [os.path.join(HERE, "csrc", "common.cuh"), os.path.join(HERE, "csrc", "plan.cuh"), os.path.join(HERE, "csrc", "sink.cuh"), os.path.join(HERE, "csrc", "prism_math.cuh"), os.path.join(HERE, "csrc", "tess_math.cuh"), os.path.join(ROOT, "include", "gravinv_b200.h")]
OUT = os.path.join(HERE, "_build", "libgravinv_b200.so")


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(f) > t for f in SRC + HDR if os.path.exists(f))


def build(force: bool = False, verbose: bool = False, out: str = OUT) -> str:
    if not force and not stale() and out == OUT:
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include")]
    cmd += os.environ.get("GI_NVCC_FLAGS", "").split()
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [s for s in SRC if os.path.exists(s)] + ["-o", out]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
