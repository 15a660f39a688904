"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference (``/root/reference``).

This module exists so that ``oracle/make_golden.py`` and the ``-m "not gpu"``
pinning tests can run the UNMODIFIED reference implementation (numpy / Cython /
numba) inside the build container and compare the oracle restatement
(``oracle/oracle_np.py``, ``oracle/csrc/*.c``) against it.  Nothing under
``gravinv3dhmc_b200/`` may import it, and nothing that runs on the GPU box may
depend on it: ``/root/reference`` does not exist there (``available()`` is then
False and callers must skip).

What is needed to import the reference on this image:

* ``numpy.float`` was removed in numpy >= 1.24 but is used at
  ``gravmag/_prism.pyx:13`` and ``gravmag/prism.py:145,309,1016`` -> alias it.
* ``matplotlib``, ``pywt``, ``vis`` are not installed / not importable -> stub
  modules (the hot path never calls into them; only the wavelet compressors
  use ``pywt`` and those cannot run here -- "parity unpinned" for that row).
* ``gravmag/_prism.pyx`` ships only as a cp37 binary -> cythonize the .pyx
  *from where it lies* into ``oracle/_ref/`` (git-ignored, travels with gpurun).

No reference source is copied into the repository.
"""
from __future__ import annotations

import glob
import importlib.util
import os
import subprocess
import sys
import sysconfig
import types

REF_ROOT = os.environ.get("GRAVINV_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
REF_BUILD = os.path.join(HERE, "_ref")

_loaded = {}


def available() -> bool:
    """True when the reference tree is mounted (build container only)."""
    return os.path.isfile(os.path.join(REF_ROOT, "inversion", "hmc.py"))


def prism_so_path():
    hits = sorted(glob.glob(os.path.join(REF_BUILD, "_prism*.so")))
    return hits[0] if hits else None


def build_ref_prism(force: bool = False) -> str:
    """Cythonize + compile the reference's ``gravmag/_prism.pyx`` into oracle/_ref/.

    Recipe (no reference build system involved): cython -> gcc -O2 -shared.
    """
    so = prism_so_path()
    if so and not force:
        return so
    if not available():
        raise RuntimeError("reference tree not mounted; cannot build oracle/_ref/_prism")
    import numpy

    os.makedirs(REF_BUILD, exist_ok=True)
    pyx = os.path.join(REF_ROOT, "gravmag", "_prism.pyx")
    c_out = os.path.join(REF_BUILD, "_prism.c")
    subprocess.check_call([sys.executable, "-m", "cython", "-3", pyx, "-o", c_out])
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    so = os.path.join(REF_BUILD, "_prism" + ext)
    inc = ["-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include()]
    subprocess.check_call(
        ["gcc", "-O2", "-fPIC", "-shared", "-fno-strict-aliasing", "-w",
         "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"] + inc + [c_out, "-o", so, "-lm"])
    return so


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _install_shims():
    import numpy

    if not hasattr(numpy, "float"):
        numpy.float = float  # _prism.pyx:13, prism.py:145,309,1016
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "matplotlib.colors",
                 "matplotlib.cm", "pywt", "mpi4py", "vis", "vis.mpl", "vis.myv"):
        if name in sys.modules:
            continue
        try:
            if name in ("pywt",):
                importlib.import_module(name)
                continue
        except Exception:
            pass
        _stub(name)
    plt = sys.modules["matplotlib.pyplot"]
    if not hasattr(plt, "switch_backend"):
        plt.switch_backend = lambda *a, **k: None
        plt.MultipleLocator = object
    sys.modules["matplotlib"].pyplot = plt
    sys.modules["matplotlib"].ticker = sys.modules["matplotlib.ticker"]
    sys.modules["vis"].mpl = sys.modules["vis.mpl"]
    sys.modules["vis"].myv = sys.modules["vis.myv"]


def load():
    """Import the reference packages; returns a namespace with the modules.

    ns.prism, ns.tesseroid, ns.potential, ns.hmc, ns.mesher, ns.utils, ns._prism
    """
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError("reference tree not mounted")
    _install_shims()
    so = build_ref_prism()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import gravmag  # noqa: F401  (reference package, from REF_ROOT)

    spec = importlib.util.spec_from_file_location("gravmag._prism", so)
    _prism = importlib.util.module_from_spec(spec)
    sys.modules["gravmag._prism"] = _prism
    spec.loader.exec_module(_prism)
    gravmag._prism = _prism

    from gravmag import prism, tesseroid, _tesseroid_numba
    import mesher
    import utils
    import constants
    from inversion import potential, hmc

    _loaded.update(prism=prism, tesseroid=tesseroid, _tesseroid_numba=_tesseroid_numba,
                   mesher=mesher, utils=utils, constants=constants,
                   potential=potential, hmc=hmc, _prism=_prism)
    return types.SimpleNamespace(**_loaded)


def load_prism_ext():
    """Load only the compiled reference ``_prism`` extension (works on the GPU box too,
    because ``oracle/_ref/`` travels with the snapshot)."""
    import numpy

    if not hasattr(numpy, "float"):
        numpy.float = float
    so = prism_so_path()
    if so is None:
        so = build_ref_prism()
    name = "_gravinv_ref_prism"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location("_prism", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


if __name__ == "__main__":
    print("reference available:", available())
    if available():
        print("built:", build_ref_prism(force="--force" in sys.argv))
