"""gravinv3dhmc_b200 -- B200-native (sm_100a CUDA behind a C ABI) implementation of the
GravInv3DHMC inversion hot path, behind the reference's Python API:

    from gravinv3dhmc_b200.inversion import potential, hmc
    from gravinv3dhmc_b200.gravmag import prism, tesseroid
    from gravinv3dhmc_b200 import mesher, utils

Only `coordinate in {"cartesian", "spherical"}` with `field="gravity"` is in scope.
"""
from . import constants  # noqa: F401

__all__ = ["constants"]
__version__ = "0.1.0"
