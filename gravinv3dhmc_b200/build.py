"""Build libgravinv_b200.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build().

Every `csrc/*.cu` is compiled to its own object (in parallel, only when it or a header changed) and
the objects are linked into one shared library; a kernel edit rebuilds one file, not ten."""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NAMES = ("assemble.cu", "leapfrog.cu", "batched.cu", "wavelet.cu", "reginv.cu", "sink.cu", "fields.cu",
         "tess_fields.cu", "fused.cu", "hostrng.cu", "peer.cu", "joint.cu")
SRC = [os.path.join(HERE, "csrc", f) for f in NAMES]
HDR = [os.path.join(HERE, "csrc", f) for f in ("common.cuh", "plan.cuh", "sink.cuh", "prism_math.cuh",
                                                "tess_math.cuh", "peer.cuh")]
HDR.append(os.path.join(ROOT, "include", "gravinv_b200.h"))
OUT = os.path.join(HERE, "_build", "libgravinv_b200.so")
OBJ = os.path.join(HERE, "_build", "obj")


def _sources():
    return [s for s in SRC if os.path.exists(s)]


def _hdr_time() -> float:
    return max([os.path.getmtime(f) for f in HDR if os.path.exists(f)] + [os.path.getmtime(__file__)])


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return _hdr_time() > t or any(os.path.getmtime(f) > t for f in _sources())


def build(force: bool = False, verbose: bool = False, out: str = OUT) -> str:
    if not force and not stale() and out == OUT:
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(OBJ, exist_ok=True)
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
             "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include")]
    extra = os.environ.get("GI_NVCC_FLAGS", "").split()
    flags += extra
    if verbose:
        flags += ["-Xptxas", "-v"]
    ht = _hdr_time()
    tag = ("_" + str(abs(hash(tuple(extra))) % 10 ** 8)) if extra else ""

    def one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + tag + ".o")
        if force or verbose or not os.path.exists(obj) or os.path.getmtime(obj) < max(ht, os.path.getmtime(src)):
            subprocess.check_call([nvcc] + flags + ["-c", src, "-o", obj])
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(one, _sources()))
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs +
                          ["-o", out, "-ldl"])
    return out


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
