"""Oracle pinning at the depth north_star states (CPU): the oracle restatement of the sampler
(oracle_np.hmc_sample / leapfrog / OracleModel on the oracle-assembled kernel) against the chains the
UNMODIFIED reference produced on configs 1-4 -- tests/golden/chains200_<case>.npz: identical
(L, accept) decisions, per-leapfrog potential and positions to 1e-9 relative.

c1_MS and c2_Smoothness run all 200 accepted samples; the other c1 / c2 cases, c3 (rejections: 491 proposals for 200 samples under MS) and c4
(256 x 72 000 kernel) run a prefix so that the CPU suite stays within minutes -- the GPU tests
(tests/test_gpu_chains200.py) compare the product with the reference over all 200."""
import numpy as np
import pytest

from tests import chains200 as c2h

PREFIX = {"c1_MS": None, "c1_Damping": 80, "c2_MS": 80, "c2_Smoothness": None,
          "c3_Damping": 60, "c3_MS": 120, "c4_Damping": 25, "c4_TV": 12}


@pytest.mark.parametrize("case", c2h.CASES)
def test_oracle_reproduces_reference_chain(case):
    from oracle import oracle_np as onp

    g = c2h.load(case)
    p = c2h.params(g)
    om, init, apr = c2h.oracle_problem(case, g)
    M = om.wm.size
    b = np.ones((M, 2))
    b[:, 0], b[:, 1] = p["bounds"]
    nprops = PREFIX[case] or g["log"].shape[0]
    reg = case.split("_")[1]
    trace = []
    out = onp.hmc_sample(om, 10 ** 6, 0, p["delta"], p["Lrange"], init, apr, b, "mandatory", 1000,
                         p["alpha"], reg, p["beta"], p["seed"], p["Sigma"], myrank=p["rank"],
                         max_proposals=nprops, trace=trace)
    idx = g["idx32"]
    U = [u for t in trace for (_, u) in t["steps"]]
    x32 = [x[idx] for t in trace for (x, _) in t["steps"]]
    c2h.compare(g, out["log"], U, x32, nprops)
    nacc = int(g["log"][:nprops, 1].sum())
    assert np.allclose(out["misfit"], g["misfit"][:nacc], rtol=1e-9, atol=0)
    if nacc >= 200:
        assert np.max(np.abs(out["models"][199] - g["model_last"])) < 1e-9 * np.max(np.abs(g["model_last"]))
    if nacc >= 100:
        assert np.max(np.abs(out["models"][99] - g["model_100"])) < 1e-9 * np.max(np.abs(g["model_100"]))
    if case == "c3_MS":
        assert 0 < nacc < nprops  # both Metropolis branches at real-data scale
