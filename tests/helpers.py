"""Shared test inputs (must mirror oracle/make_golden.py so fixtures and live runs agree)."""
import numpy as np


def small_prism_setup():
    xs = np.linspace(0, 400, 4)
    ys = np.linspace(0, 600, 6)
    X, Y = np.meshgrid(xs, ys)
    xp = np.concatenate([X.ravel(), [50.0, 150.0, 333.3, 410.0, -20.0, 200.0]])
    yp = np.concatenate([Y.ravel(), [50.0, 250.0, 123.4, 610.0, -30.0, 300.0]])
    zp = np.concatenate([np.zeros(24), [-1.0, -50.0, -10.0, -5.0, -0.5, -150.0]])
    return xp, yp, zp


def synthetic_topo(x1, x2, y1, y2, amp, base, n=9):
    xs = np.linspace(x1, x2, n)
    ys = np.linspace(y1, y2, n)
    X, Y = np.meshgrid(xs, ys)
    H = base + amp * np.sin(2.1 * (X - x1) / (x2 - x1) + 0.3) * np.cos(1.7 * (Y - y1) / (y2 - y1))
    return X.ravel(), Y.ravel(), H.ravel()


def chain_params(g, name):
    alpha, beta, delta, Sigma, L0, L1, seed, nsamples, lo, hi = g[f"chain_{name}_params"]
    reg = name if name in ("Damping", "MS", "Smoothness", "TV") else "Damping"
    return dict(alpha=float(alpha), beta=float(beta), delta=float(delta), Sigma=float(Sigma),
                Lrange=[int(L0), int(L1)], seed=int(seed), nsamples=int(nsamples), lo=float(lo),
                hi=float(hi), reg=reg,
                constraint="logarithmic" if name == "log" else "mandatory",
                init=0.3 if name == "log" else 0.001, fixed=(name == "fixed"))


def chain_from_golden(g, name, runner="oracle"):
    """Re-run one golden chain with the oracle; returns the same dict layout make_golden stores."""
    from oracle import oracle_np as onp

    p = chain_params(g, name)
    dobs = g["fixed_dobs"] if p["fixed"] else g["small_dobs"]
    model = onp.OracleModel(g["small_Aw"], g["small_wm"], dobs, tuple(g["small_mshape"]),
                            fixed=p["fixed"], grav_fix=g["fixed_gravfix"] if p["fixed"] else None)
    M = g["small_wm"].size
    b = np.ones((M, 2))
    b[:, 0], b[:, 1] = p["lo"], p["hi"]
    nprops = g[f"chain_{name}_prop_log"].shape[0]
    trace = []
    out = onp.hmc_sample(model, p["nsamples"], 0, p["delta"], p["Lrange"], np.ones(M) * p["init"],
                         np.ones(M) * p["init"], b, p["constraint"], 1000, p["alpha"], p["reg"],
                         p["beta"], p["seed"], p["Sigma"], max_proposals=nprops, trace=trace)
    xs, Us, log = [], [], []
    for t in trace:
        for (x, U) in t["steps"]:
            xs.append(x)
            Us.append(U)
        log.append((t["L"], int(t["accept"])))
    return dict(steps_x=np.array(xs), steps_U=np.array(Us), prop_log=np.array(log),
                misfit=out["misfit"], models=out["models"])
