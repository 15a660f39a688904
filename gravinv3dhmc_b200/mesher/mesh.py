"""Mesh bookkeeping: flat index -> cell bounds, topography masks, axis coordinates.

Host-side mirror of the reference's mesher/mesh.py (PrismMesh :126-223/:229-270, carvetopo
:301-394, get_xs/ys/zs :396-445; PrismMeshSegment :561-686, carvetopo :717-797; TesseroidMesh
:518-558; TesseroidMeshSegment :914-955).  Same constructor arguments, attributes (`bounds`,
`dims`, `shape`, `size`, `mask`, `props`, `zdown`) and iteration protocol, but the per-cell Python
loop is replaced by `bounds_table()`, which produces the [M, 6] table of ACTIVE cells that the CUDA
assembly kernels consume.  Index, bounds and mask values are bit-identical to the reference's:
every coordinate is formed by the same scalar double-precision expression (per axis value, then
broadcast), never by a re-associated vectorised formula.
"""
from __future__ import annotations

import copy as cp

import numpy as np

from .geometry import Prism, Tesseroid


class _MeshBase:
    celltype = Prism

    # ---- set by subclasses: self.bounds, self.dims, self.shape, self.size ----
    def _init_common(self, props):
        self.props = {} if props is None else props
        self.i = 0
        self.mask = []      # flat indices of masked cells, ascending (reference: python list)
        self.zdown = True
        self._mask_cache = None

    def __len__(self):
        return self.size

    # per-axis lower/upper edges, each value computed by the reference's scalar expression
    def _axis_edges(self):
        nz, ny, nx = self.shape
        dx, dy = self.dims[0], self.dims[1]
        b = self.bounds
        x1 = np.array([b[0] + dx * i for i in range(nx)], dtype=np.float64)
        x2 = np.array([(b[0] + dx * i) + dx for i in range(nx)], dtype=np.float64)
        y1 = np.array([b[2] + dy * j for j in range(ny)], dtype=np.float64)
        y2 = np.array([(b[2] + dy * j) + dy for j in range(ny)], dtype=np.float64)
        z1 = np.empty(nz)
        z2 = np.empty(nz)
        for k in range(nz):
            z1[k], z2[k] = self._layer(k)
        return x1, x2, y1, y2, z1, z2

    def node_axes(self):
        """(xn, yn, zn) node coordinates per axis when neighbouring cells share their edges bit for
        bit (upper edge of cell i == lower edge of cell i+1 as doubles on every axis), else None.
        When they do, a corner of the mesh is one point for all the cells around it and the
        assembly kernel may evaluate it once (gi_prism_gz_assemble_grid)."""
        x1, x2, y1, y2, z1, z2 = self._axis_edges()
        for lo, hi in ((x1, x2), (y1, y2), (z1, z2)):
            if not np.array_equal(hi[:-1], lo[1:]):
                return None
        return (np.append(x1, x2[-1]), np.append(y1, y2[-1]), np.append(z1, z2[-1]))

    def column_map(self):
        """int32 [size]: column of G for every cell (-1 = masked), or None when nothing is masked"""
        if not len(self.mask):
            return None
        m = self._mask_array()
        cm = np.full(self.size, -1, dtype=np.int32)
        cm[~m] = np.arange(int((~m).sum()), dtype=np.int32)
        return cm

    def _mask_array(self):
        """boolean [size] array of masked cells (cached against the mask list's length)"""
        key = len(self.mask)
        if self._mask_cache is None or self._mask_cache[0] != key:
            m = np.zeros(self.size, dtype=bool)
            if key:
                m[np.asarray(self.mask, dtype=np.int64)] = True
            self._mask_cache = (key, m)
        return self._mask_cache[1]

    def active_indices(self):
        """flat indices of the un-masked cells, ascending (= column order of G)"""
        return np.flatnonzero(~self._mask_array())

    def bounds_table(self, active_only=True):
        """[M, 6] float64 table (x1,x2,y1,y2,z1,z2 | w,e,s,n,top,bottom) in mesh order
        (x fastest, then y, then z), masked cells removed when `active_only`."""
        nz, ny, nx = self.shape
        x1, x2, y1, y2, z1, z2 = self._axis_edges()
        tab = np.empty((nz, ny, nx, 6), dtype=np.float64)
        tab[..., 0] = x1[None, None, :]
        tab[..., 1] = x2[None, None, :]
        tab[..., 2] = y1[None, :, None]
        tab[..., 3] = y2[None, :, None]
        tab[..., 4] = z1[:, None, None]
        tab[..., 5] = z2[:, None, None]
        tab = tab.reshape(-1, 6)
        if active_only and len(self.mask):
            tab = tab[~self._mask_array()]
        return np.ascontiguousarray(tab)

    def __getitem__(self, index):
        if index >= self.size or index < -self.size:
            raise IndexError("mesh index out of range")
        if index < 0:
            index = self.size + index
        if self._mask_array()[index]:
            return None
        nz, ny, nx = self.shape
        k = index // (nx * ny)
        j = (index - k * (nx * ny)) // nx
        i = index - k * (nx * ny) - j * nx
        x1 = self.bounds[0] + self.dims[0] * i
        x2 = x1 + self.dims[0]
        y1 = self.bounds[2] + self.dims[1] * j
        y2 = y1 + self.dims[1]
        z1, z2 = self._layer(k)
        props = dict([p, self.props[p][index]] for p in self.props)
        return self.celltype(x1, x2, y1, y2, z1, z2, props=props)

    def __iter__(self):
        self.i = 0
        return self

    def __next__(self):
        if self.i >= self.size:
            raise StopIteration
        cell = self.__getitem__(self.i)
        self.i += 1
        return cell

    def addprop(self, prop, values):
        self.props[prop] = values

    def copy(self):
        return cp.deepcopy(self)

    def get_xs(self):
        x1, x2 = self.bounds[0], self.bounds[1]
        dx = self.dims[0]
        xs = np.arange(x1, x2 + dx, dx)
        return xs[:-1] if xs.size > self.shape[2] + 1 else xs

    def get_ys(self):
        y1, y2 = self.bounds[2], self.bounds[3]
        dy = self.dims[1]
        ys = np.arange(y1, y2 + dy, dy)
        return ys[:-1] if ys.size > self.shape[1] + 1 else ys

    def get_layer(self, i):
        nz, ny, nx = self.shape
        if i >= nz or i < 0:
            raise IndexError("Layer index %d is out of range." % (i))
        return [self.__getitem__(p) for p in range(i * nx * ny, (i + 1) * nx * ny)]

    def layers(self):
        for i in range(self.shape[0]):
            yield self.get_layer(i)

    # ---- topography ----
    def _carve(self, x, y, height, below, zc, method, write_interp):
        import scipy.interpolate

        nz, ny, nx = self.shape
        x1, x2, y1, y2 = self.bounds[:4]
        dx, dy = self.dims[0], self.dims[1]
        xc = np.arange(x1, x2, dx) + 0.5 * dx
        if len(xc) > nx:
            xc = xc[:-1]
        yc = np.arange(y1, y2, dy) + 0.5 * dy
        if len(yc) > ny:
            yc = yc[:-1]
        if len(zc) > nz:
            zc = zc[:-1]
        XC, YC = np.meshgrid(xc, yc)
        topo = scipy.interpolate.griddata((x, y), height, (XC, YC), method=method).ravel()
        if self.zdown:
            topo = -1 * topo
        if write_interp:  # the reference always writes this file into the CWD (mesh.py:372,775)
            np.savetxt("carve_topo_interp.txt", np.c_[XC.ravel(), YC.ravel(), topo], fmt="%.8f",
                       delimiter=" ")
        masked = np.asarray(topo.mask if np.ma.isMA(topo) else np.zeros(topo.shape, dtype=bool))
        t = np.asarray(topo, dtype=np.float64)[None, :]
        zc = np.asarray(zc, dtype=np.float64)[:, None]
        # comparisons with NaN are False, like the reference's scalar tests
        if below:
            hit = (zc > t) if self.zdown else (zc < t)
        else:
            hit = (zc < t) if self.zdown else (zc > t)
        hit = hit | masked[None, :]
        self.mask.extend(int(c) for c in np.flatnonzero(hit.ravel()))
        self._mask_cache = None
        return self.mask


class PrismMesh(_MeshBase):
    """Regular (optionally geometrically stretched in z) mesh of right rectangular prisms.

    bounds = [xmin, xmax, ymin, ymax, zmin, zmax]; spacing = (dz, dy, dx); ratio = growth of dz
    with depth (mesh.py:166-223)."""

    celltype = Prism

    def __init__(self, bounds, spacing, ratio=1, props=None, verbose=False):
        dz, dy, dx = spacing
        x1, x2, y1, y2, z1, z2 = bounds
        self.dims = (dx, dy, dz)
        self.ratio = ratio
        nx = int(np.ceil((x2 - x1) / dx))
        ny = int(np.ceil((y2 - y1) / dy))
        if ratio == 1:
            nz = int(np.ceil((z2 - z1) / dz))
            bounds_big = x1, x1 + nx * dx, y1, y1 + ny * dy, z1, z1 + nz * dz
        else:
            n = 1
            while True:
                depth = z1 + dz * (1 - ratio ** n) / (1 - ratio)
                if depth < z2 and (z2 - depth) > dz:
                    n += 1
                else:
                    break
            nz = int(n)
            bounds_big = x1, x1 + nx * dx, y1, y1 + ny * dy, z1, z2
        if verbose:
            print("grid with new boundaries: {}".format(bounds_big))
        self.bounds = bounds_big
        self.shape = tuple(int(i) for i in (nz, ny, nx))
        self.size = int(nx * ny * nz)
        self._init_common(props)

    def _layer(self, k):
        nz = self.shape[0]
        if self.ratio == 1:  # mesh.py:246-256
            z1 = self.bounds[4] + self.dims[2] * k
            z2 = z1 + self.dims[2] if k < nz - 1 else self.bounds[5]
        else:  # mesh.py:258-267
            z2 = self.bounds[4] + self.dims[2] * (1 - self.ratio ** (k + 1)) / (1 - self.ratio)
            z1 = z2 - self.dims[2] * self.ratio ** k
            if k == nz - 1:
                z2 = self.bounds[5]
        return z1, z2

    def get_zs(self):
        z1, z2 = self.bounds[4], self.bounds[5]
        dz = self.dims[2]
        nz = self.shape[0]
        if self.ratio == 1:
            zs = np.arange(z1, z2 + dz, dz)
        else:
            zs = np.zeros(nz + 1)
            for k in range(nz):
                bottom = self.bounds[4] + dz * (1 - self.ratio ** (k + 1)) / (1 - self.ratio)
                zs[k] = bottom - dz * self.ratio ** k
            zs[nz] = z2
        return zs[:-1] if zs.size > nz + 1 else zs

    def carvetopo(self, x, y, height, below=False, write_interp=True):
        """Mask cells above the topography: cell CENTRES vs cubic interpolation (mesh.py:301-394)."""
        nz = self.shape[0]
        z1, z2 = self.bounds[4], self.bounds[5]
        dz = self.dims[2]
        if self.ratio == 1:
            zc = np.arange(z1, z2, dz) + 0.5 * dz
        else:
            zc = np.zeros(nz)
            bottom = None
            for k in range(0, nz - 1):
                bottom = self.bounds[4] + dz * (1 - self.ratio ** (k + 1)) / (1 - self.ratio)
                zc[k] = bottom - 0.5 * dz * self.ratio ** k
            zc[nz - 1] = bottom + 0.5 * (z2 - bottom)
        return self._carve(x, y, height, below, zc, "cubic", write_interp)


class TesseroidMesh(PrismMesh):
    """bounds = [w, e, s, n, top, bottom] (degrees, metres); spacing = (dr, dlat, dlon) with a
    NEGATIVE dr (heights decrease with the layer index); mesh.py:518-558."""

    celltype = Tesseroid

    def __init__(self, bounds, spacing, ratio=1, props=None, verbose=False):
        super().__init__(bounds, spacing, ratio, props, verbose)
        self.zdown = False
        self.dump = None


class PrismMeshSegment(_MeshBase):
    """Prism mesh whose z spacing is piecewise constant: spacing = ([dz1, dz2, ...], dy, dx),
    divisionsection = [z0, z1, ..., zn] (mesh.py:601-645)."""

    celltype = Prism

    def __init__(self, bounds, spacing, divisionsection, props=None, verbose=False):
        x1, x2, y1, y2, z1, z2 = bounds
        dzlist, dy, dx = spacing
        self.dims = (dx, dy, dzlist)
        self.segment = len(dzlist)
        self.divisionsection = divisionsection
        nx = int(np.ceil((x2 - x1) / dx))
        ny = int(np.ceil((y2 - y1) / dy))
        nz = 0
        nzlist = np.zeros(self.segment)
        nzsumlist = np.zeros(self.segment)
        for i in range(self.segment):
            nzlist[i] = int(np.ceil((divisionsection[i + 1] - divisionsection[i]) / dzlist[i]))
            nz = nz + nzlist[i]
            nzsumlist[i] = nz
        bounds_big = (x1, x1 + nx * dx, y1, y1 + ny * dy, z1,
                      self.divisionsection[-2] + nzlist[-1] * dzlist[-1])
        if verbose:
            print("Uniform Segment grid with new boundaries: {}".format(bounds_big))
        self.nzlist = nzlist
        self.nzsumlist = nzsumlist
        self.bounds = bounds_big
        self.shape = tuple(int(i) for i in (nz, ny, nx))
        self.size = int(nx * ny * nz)
        self._init_common(props)

    def _layer(self, k):
        # mesh.py:669-683 (nzsumlist holds floats; k - nzsumlist[...] is a float)
        kloc = None
        for iseg in range(self.segment):
            if k < self.nzsumlist[iseg]:
                kloc = iseg
                break
        if kloc == 0:
            z1 = self.bounds[4] + self.dims[2][kloc] * k
        else:
            z1 = self.divisionsection[kloc] + self.dims[2][kloc] * (k - self.nzsumlist[kloc - 1])
        z2 = z1 + self.dims[2][kloc]
        return float(z1), float(z2)

    def _segment_tops(self):
        zs = []
        for iseg in range(self.segment):
            zs.extend(list(np.arange(self.divisionsection[iseg], self.divisionsection[iseg + 1],
                                     self.dims[2][iseg])))
        return zs

    def get_zs(self):
        zs = self._segment_tops()
        zs.append(self.bounds[5])
        zs = np.array(zs)
        return zs[:-1] if zs.size > self.shape[0] + 1 else zs

    def carvetopo(self, x, y, height, below=False, write_interp=True):
        """Mask cells above the topography: cell TOPS vs nearest-neighbour interpolation
        (mesh.py:717-797)."""
        return self._carve(x, y, height, below, np.array(self._segment_tops()), "nearest",
                           write_interp)


class TesseroidMeshSegment(PrismMeshSegment):
    """mesh.py:914-955"""

    celltype = Tesseroid

    def __init__(self, bounds, spacing, divisionsection, props=None, verbose=False):
        super().__init__(bounds, spacing, divisionsection, props, verbose)
        self.zdown = False
        self.dump = None
