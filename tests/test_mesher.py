"""Product mesher vs golden vectors of the reference's mesher (bit-exact) -- CPU only."""
import numpy as np
import pytest

from gravinv3dhmc_b200 import mesher, utils
from tests.test_oracle_pinning import MESH_CASES

CLS = {"prism_uniform": mesher.PrismMesh, "prism_uniform_nondiv": mesher.PrismMesh,
       "prism_ratio": mesher.PrismMesh, "prism_segment": mesher.PrismMeshSegment,
       "tess_uniform": mesher.TesseroidMesh, "tess_segment": mesher.TesseroidMeshSegment}


def make(name):
    kw = dict(MESH_CASES[name])
    kw.pop("zdown", None)
    args = [kw.pop("bounds"), kw.pop("spacing")]
    if "divisionsection" in kw:
        args.append(kw.pop("divisionsection"))
    return CLS[name](*args, **kw)


@pytest.mark.parametrize("name", list(MESH_CASES))
def test_mesh_bit_exact(golden, name):
    g = golden["meshes"]
    m = make(name)
    assert m.shape == tuple(g[name + "_shape"])
    assert m.size == g[name + "_table"].shape[0]
    assert np.array_equal(np.array(m.bounds, dtype=np.float64), g[name + "_bounds"])
    assert np.array_equal(m.bounds_table(), g[name + "_table"])
    assert np.array_equal(m.get_xs(), g[name + "_xs"])
    assert np.array_equal(m.get_ys(), g[name + "_ys"])
    assert np.array_equal(m.get_zs(), g[name + "_zs"])
    # iteration protocol hands out the same cells
    for idx in (0, m.size // 2, m.size - 1, -1):
        assert np.array_equal(np.array(m[idx].get_bounds()), g[name + "_table"][idx])
    assert len(list(iter(m))) == m.size


def test_carve_masks_bit_exact(golden, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    g = golden["meshes"]
    for name, key in (("prism_uniform", "carve_prism"), ("prism_segment", "carve_prismseg"),
                      ("tess_segment", "carve_tessseg")):
        m = make(name)
        t = g[key + "_topo"]
        mask = m.carvetopo(t[:, 0], t[:, 1], t[:, 2])
        assert isinstance(mask, list)
        assert np.array_equal(np.array(mask), g[key + "_mask"])
        assert (tmp_path / "carve_topo_interp.txt").exists()
        assert m[int(mask[0])] is None
    gt = g["carve_tessseg_table"]
    live = ~np.isnan(gt[:, 0])
    assert np.array_equal(m.bounds_table(), gt[live])
    assert np.array_equal(m.active_indices(), np.flatnonzero(live))
    rho = np.arange(m.size, dtype=np.float64) * 0.5
    rc = utils.rho2carve(rho, m.mask)
    assert np.array_equal(rc, g["rho2carve_out"])
    full = np.full(m.size, -7.0)
    out = utils.carve2rho(rc + 1.0, full, m.mask)
    assert np.array_equal(out, g["carve2rho_out"])
    assert np.array_equal(full, out)  # in place, like the reference
