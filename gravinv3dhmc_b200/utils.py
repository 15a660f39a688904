"""Model-vector compaction for topography-carved meshes (reference utils.py:714-749) and the
direction helpers of the magnetic fields (utils.py:420-474)."""
import numpy as np


def _keep(n, mask):
    keep = np.ones(n, dtype=bool)
    m = np.asarray(mask, dtype=np.int64)
    if m.size:
        keep[m] = False
    return keep


def rho2carve(rho, mask):
    """Drop the masked cells of a full-grid vector (utils.py:714-727)."""
    rho = np.asarray(rho)
    return rho[_keep(rho.shape[0], mask)].copy()


def carve2rho(rhocarve, rho, mask):
    """Scatter a carved vector back into the full-grid vector `rho` (updated in place, as the
    reference does) and return a copy (utils.py:729-749)."""
    keep = _keep(rho.shape[0], mask)
    rho[keep] = np.asarray(rhocarve)[: int(keep.sum())]
    return rho.copy()


def dircos(inc, dec):
    """unit vector [x, y, z] (x North, y East, z Down) of an inclination / declination in degrees
    (utils.py:448-474)"""
    d2r = np.pi / 180.
    return [np.cos(d2r * inc) * np.cos(d2r * dec), np.cos(d2r * inc) * np.sin(d2r * dec),
            np.sin(d2r * inc)]


def ang2vec(intensity, inc, dec):
    """vector(s) of the given intensity along (inc, dec) (utils.py:420-445)"""
    return np.transpose([intensity * i for i in dircos(inc, dec)])
