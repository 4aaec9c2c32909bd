"""tests/golden/*.npz: committed outputs of short seeded scan streams (tests/golden/make_golden.py).

CPU: the oracle still reproduces them, and the synthetic generator still produces the inputs they
were made from.  GPU: the CUDA path, through the C-ABI, reproduces the same committed data."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import make_golden as mg  # noqa: E402
from parity_utils import EXACT_LAYERS, compare_layer, run_pair  # noqa: E402


def _load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_inputs_of_the_fixture_are_reproducible(name):
    wl, n_scans, _ = mg.case_config(name)
    want = bytes(_load(name)["input_sha256"]).decode()
    assert mg.input_digest(wl, n_scans) == want, "fastdem_b200/synthetic.py no longer generates the fixture's scans"


@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_oracle_reproduces_the_fixture(name):
    fx = _load(name)
    now = mg.run_oracle(name)
    assert sorted(k for k in now if k.startswith("layer:")) == sorted(k for k in fx.files if k.startswith("layer:"))
    assert np.array_equal(now["stats"], fx["stats"])
    assert np.array_equal(now["geometry"], fx["geometry"])
    assert np.array_equal(now["position"], fx["position"])
    for key in fx.files:
        if not key.startswith("layer:"):
            continue
        got, want = now[key], fx[key]
        assert np.array_equal(np.isnan(got), np.isnan(want)), key
        ok = ~np.isnan(want)
        if key[6:] in EXACT_LAYERS:
            assert np.array_equal(got[ok].view(np.uint32), want[ok].view(np.uint32)), key
        else:   # same compiler flags, same libm: normally bit-identical; the bound is the north-star tolerance
            assert np.allclose(got[ok], want[ok], rtol=1e-6, atol=1e-7), key


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_cuda_path_reproduces_the_fixture(fdem, name):
    fx = _load(name)
    wl_name, n_scans, over = mg.CASES[name]
    from fastdem_b200 import synthetic as syn
    wl = syn.WORKLOADS[wl_name]
    cfg = wl.config()
    for k, v in over.items():
        setattr(cfg, k, v)
    gmap, omap, gdem, odem, gs, os_ = run_pair(fdem, wl, n_scans, cfg)
    assert sorted(gmap.getLayers()) == sorted(k[6:] for k in fx.files if k.startswith("layer:"))
    assert [[int(s.integrated), s.n_kept, s.n_cells, s.n_voxels] for s in gs] == fx["stats"].tolist()
    gg = gmap.geometry()
    assert [gg.rows, gg.cols, gg.start_index[0], gg.start_index[1]] == fx["geometry"].tolist()
    assert [gg.position[0], gg.position[1]] == fx["position"].tolist()
    for key in fx.files:
        if key.startswith("layer:"):
            compare_layer(key[6:], gmap.get(key[6:]), fx[key])
