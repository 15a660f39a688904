#!/usr/bin/env python
"""Run one of the reference's shipped examples through gravinv3dhmc_b200, driven by the example's own
`SetPMTS.txt` -- the drop-in proof for the path BASELINE.json names:

    python tools/run_example.py <reference>/example/uniformgrid 0          # main_uniform.py 0
    python tools/run_example.py <reference>/example/segmentgrid 0          # main_seg.py 0
    python tools/run_example.py <reference>/example/realdata 0             # main_real.py 0
    python tools/run_example.py <reference>/example/global 0 --chains 2    # mpiexec -n 2 main_global.py 0

What the reference's drivers do (example/uniformgrid/main_uniform.py:98-119 and its three siblings):
every line of `SetPMTS.txt` is a Python dict literal (`eval(line)` there, `ast.literal_eval` here) with
the keys `set, test, rhomin, rhomax, mspacing, Lrange, delta, Sigma, RegulFactor, regularization, beta,
nsamples`; the line number is the first command-line argument; `mpiexec -n K` runs K independent chains
whose rank seeds the RNG (seed 100 + rank) and names the output folder `result/<set><test>_chain<rank>`.
The model geometry, the data files and the start / prior models are hard-wired per example in
`main_*.py`; they are restated in `EXAMPLES` below with the line they come from.

Two quirks of the shipped files are honoured: `main_global.py:22-28` reorders its `mspacing`
`[dlon, dlat, dr]` to `(dr, dlat, dlon)`; `example/segmentgrid/SetPMTS.txt` ships
`[100, 100, [100, 200, 300]]` although `main_seg.py:37` passes it on unchanged to a mesh that expects the
layer list FIRST (`mesher/mesh.py:603`; the shipped log shows `[100, 200, 300]` was what ran) -- a
list in the last position is therefore moved to the front.

Chains: `--chains 1` is `hmc.HMCSample(myrank=0)`; `--chains K` runs ranks 0..K-1 as ONE batched
device loop (`batched.HMCSampleBatch`), each chain bit-compatible with the reference process of that rank.
"""
from __future__ import annotations

import argparse
import ast
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KEYS = ("set", "test", "rhomin", "rhomax", "mspacing", "Lrange", "delta", "Sigma", "RegulFactor",
        "regularization", "beta", "nsamples")

# per-example constants of the shipped drivers
EXAMPLES = {
    # example/uniformgrid/main_uniform.py:26-40
    "uniformgrid": dict(mrange=(0, 2000, 0, 3000, 0, 1000), data="modeldata/{set}_gz_noise.txt",
                        kw=dict(coordinate="cartesian"), wavelet="3D", init=0.001, apr=0.001),
    # example/segmentgrid/main_seg.py:27-39
    "segmentgrid": dict(mrange=(0, 2000, 0, 3000, 0, 2100), data="modeldata/{set}_gz_noise.txt",
                        kw=dict(coordinate="cartesian", mseg=True, mdivisionsection=[0, 300, 900, 2100]),
                        wavelet="3D", init=0.001, apr=0.001),
    # example/realdata/main_real.py:25-74
    "realdata": dict(mrange=(106.5, 118.5, 16, 28, 2000, -60000), data="data/gravinv_12d05d.dat",
                     kw=dict(coordinate="spherical", mseg=True, mdivisionsection=[2000, -5000, -15000, -60000],
                             fixed=True),
                     fix="data/grasea_12d05d.dat", topo="data/topo_12d05d.dat",
                     apr_file="data/SC_ApriorModel.txt", wavelet=False, init=0.01),
    # example/global/main_global.py:25-36
    "global": dict(mrange=(-180, 180, -90, 90, 0, -3000000), data="modeldata/{set}_gz_noise.txt",
                   kw=dict(coordinate="spherical"), wavelet=False, init=0.001, apr=0.001, reorder=True),
}


def parse_setpmts(path):
    """every non-empty line of SetPMTS.txt as a dict (main_uniform.py:98-102, without `eval`)"""
    out = []
    with open(path, "r") as f:
        for line in f:
            if line.strip():
                d = ast.literal_eval(line.strip())
                missing = [k for k in KEYS if k not in d]
                if missing:
                    raise ValueError("SetPMTS.txt line %d lacks %s" % (len(out), missing))
                out.append(d)
    return out


def example_kind(example_dir, kind=None):
    kind = kind or os.path.basename(os.path.normpath(example_dir))
    if kind not in EXAMPLES:
        raise ValueError("unknown example %r: pass --kind from %s" % (kind, sorted(EXAMPLES)))
    return kind


def model_spacing(kind, mspacing):
    """the `mspacing` the example's driver hands to GravMagModule"""
    sp = list(mspacing)
    if EXAMPLES[kind].get("reorder"):      # main_global.py:22-28: [dlon, dlat, dr] -> (dr, dlat, dlon)
        return (sp[2], sp[1], sp[0])
    if isinstance(sp[2], (list, tuple)) and not isinstance(sp[0], (list, tuple)):
        return (list(sp[2]), sp[1], sp[0])  # segmentgrid/SetPMTS.txt as shipped (see the module docstring)
    return tuple(sp)


def build(example_dir, pm, kind, wavelet="shipped", verbose=True):
    """(model, dobs, initial_model, aprior_model, boundaries) exactly as the example's main() sets them up"""
    from gravinv3dhmc_b200 import utils
    from gravinv3dhmc_b200.inversion import potential

    ex = EXAMPLES[kind]
    p = lambda rel: os.path.join(example_dir, rel.format(set=pm["set"]))
    xo, yo, ho, dobs = np.loadtxt(p(ex["data"]), usecols=[0, 1, 2, 3], unpack=True)
    kw = dict(ex["kw"])
    if "fix" in ex:
        kw["grav_fix"] = np.loadtxt(p(ex["fix"]), usecols=[2], unpack=True)
    if "topo" in ex:
        kw["mtopo"] = tuple(np.loadtxt(p(ex["topo"]), usecols=[0, 1, 2], unpack=True))
    wv = ex["wavelet"] if wavelet == "shipped" else (False if wavelet in ("none", "False") else wavelet)
    model = potential.GravMagModule(dobs, ex["mrange"], model_spacing(kind, pm["mspacing"]), (xo, yo, ho),
                                    njobs=5, field="gravity", wavelet=wv, verbose=verbose, **kw)
    ncell = int(np.prod(model.mshape))
    if "apr_file" in ex:    # main_real.py:68-74: constant start model, prior from file, both carved
        init = utils.rho2carve(np.ones(ncell) * ex["init"], model.mask)
        apr = utils.rho2carve(np.loadtxt(p(ex["apr_file"]), usecols=[3], unpack=True), model.mask)
    else:
        init, apr = np.ones(ncell) * ex["init"], np.ones(ncell) * ex["apr"]
    b = np.ones((init.shape[0], 2))
    b[:, 0], b[:, 1] = pm["rhomin"], pm["rhomax"]
    return model, dobs, init, apr, b


def run(example_dir, line, kind=None, nsamples=None, chains=1, out=None, wavelet="shipped", quiet=False,
        output="text"):
    from gravinv3dhmc_b200.inversion import batched, hmc

    kind = example_kind(example_dir, kind)
    pm = parse_setpmts(os.path.join(example_dir, "SetPMTS.txt"))[line]
    nsamples = pm["nsamples"] if nsamples is None else nsamples
    t0 = time.time()
    model, dobs, init, apr, b = build(example_dir, pm, kind, wavelet, verbose=not quiet)
    t_model = time.time() - t0
    out = out or os.path.join(example_dir, "result")
    os.makedirs(out, exist_ok=True)
    save_folder = os.path.join(out, str(pm["set"]) + str(pm["test"]) + "_chain")
    # main_uniform.py:52-76: seed 100, ndraws 0, "Fixed" regularisation factor, mandatory constraint
    args = (pm["delta"], pm["Lrange"], init, apr, b, "mandatory", 1000, dobs, "Fixed", 0.8,
            pm["RegulFactor"], pm["regularization"], pm["beta"], 100, pm["Sigma"])
    t0 = time.time()
    if chains == 1:
        ch = hmc.setup_chain(model, *args, myrank=0, save_folder=save_folder, quiet=quiet)
        ch.output = output
        ch.sample(nsamples, 0)
        proposals = [ch.proposals]
    else:
        if model.wavelet:
            raise SystemExit("the wavelet-compressed forward is a single-chain path: use --chains 1 or --wavelet none")
        bt = batched.HMCBatch(model, chains, args[0], args[1], init, apr, b, "mandatory", 1000, dobs, args[10],
                              args[11], args[12], 100, args[14], save_folder=save_folder, quiet=quiet)
        bt.output = output
        bt.stream(nsamples, 0)
        proposals = bt.proposals
        bt.close()
    t_chain = time.time() - t0
    return dict(kind=kind, params=pm, model_seconds=t_model, chain_seconds=t_chain, save_folder=save_folder,
                proposals=proposals, shape=tuple(model.mshape), voxels=model.M, observations=model.n_total)


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("example_dir")
    ap.add_argument("line", type=int, help="line of SetPMTS.txt (the reference's sys.argv[1])")
    ap.add_argument("--kind", choices=sorted(EXAMPLES), help="default: the directory's name")
    ap.add_argument("--nsamples", type=int, help="override SetPMTS's nsamples")
    ap.add_argument("--chains", type=int, default=1, help="`mpiexec -n K`: ranks 0..K-1 as one batch")
    ap.add_argument("--out", help="result directory (default <example_dir>/result)")
    ap.add_argument("--wavelet", default="shipped", help="shipped | none | 1D | 3D")
    ap.add_argument("--output", default="text", choices=["text", "binary", "none"])
    ap.add_argument("--quiet", action="store_true")
    a = ap.parse_args()
    r = run(a.example_dir, a.line, a.kind, a.nsamples, a.chains, a.out, a.wavelet, a.quiet, a.output)
    acc = [sum(1 for _, ok in p if ok) for p in r["proposals"]]
    print("%s %s%s: grid %s (%d voxels) x %d observations; kernel + weighting %.2f s; %d chain(s), "
          "%s accepted of %s proposals in %.2f s -> %s<rank>/"
          % (r["kind"], r["params"]["set"], r["params"]["test"], r["shape"], r["voxels"], r["observations"],
             r["model_seconds"], len(acc), acc, [len(p) for p in r["proposals"]], r["chain_seconds"],
             r["save_folder"]))


if __name__ == "__main__":
    main()
