"""Forward-only tesseroid modelling on the GPU (synthetic-data generation).

Mirror of the reference's gravmag/tesseroidforward.py (`potential, gx, gy, gz, gxx, gxy, gxz, gyy,
gyz, gzz`, :236-788, over gravmag/_tesseroid_numba_forward.py): same arguments, the result array
only, the reference's scale factors (`result *= SI2MGAL*G` etc.: one multiplication by the product;
gy with G here, unlike gravmag/tesseroid.py:416-417).  The kernel is assembled in observation chunks
by `gi_tess_field_assemble` / the gz kernel's engine and contracted with the densities on the device,
so the footprint stays bounded however large the model is.
"""
from __future__ import annotations

from . import tesseroid as _t

RATIO_V, RATIO_G, RATIO_GG, STACK_SIZE = _t.RATIO_V, _t.RATIO_G, _t.RATIO_GG, _t.STACK_SIZE


def _make(field, ratio0, lines):
    def fn(lon, lat, height, model, dens=None, ratio=ratio0, njobs=1, pool=None):
        return _t._field(field, lon, lat, height, model, dens, ratio, njobs, pool, False, forward=True)

    fn.__name__ = field
    fn.__doc__ = "{} of a tesseroid model (gravmag/tesseroidforward.py:{}).".format(field, lines)
    return fn


potential = _make("potential", RATIO_V, "236-288")
gx = _make("gx", RATIO_G, "291-343")
gy = _make("gy", RATIO_G, "346-398")
gz = _make("gz", RATIO_G, "401-458")
gxx = _make("gxx", RATIO_GG, "461-513")
gxy = _make("gxy", RATIO_GG, "516-568")
gxz = _make("gxz", RATIO_GG, "571-623")
gyy = _make("gyy", RATIO_GG, "626-678")
gyz = _make("gyz", RATIO_GG, "681-733")
gzz = _make("gzz", RATIO_GG, "736-788")
