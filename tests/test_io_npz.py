"""Map checkpoint IO (fastdem_b200.io_npz) vs the oracle's restatement of io_npz.cpp and vs
numpy's own reader.  Cases follow the reference's tests/test_map_io.cpp.

CPU tests drive the product reader/writer through a host stand-in for the map (same duck-typed
surface the device ElevationMap offers); the gpu-marked tests use the real device map."""
import types

import numpy as np
import pytest

import oracle_binding as ob
from fastdem_b200 import io_npz


class HostMap:
    """Host stand-in with the ElevationMap methods io_npz uses (geometry rule = the oracle's)."""

    def __init__(self, width=0.0, height=0.0, resolution=0.0, frame_id=""):
        self._frame = frame_id
        self._layers = {}
        self._order = []
        if resolution > 0:
            self.setGeometry(width, height, resolution)

    def setGeometry(self, width, height, resolution):
        om = ob.OracleMap(width, height, resolution)
        g = om.geometry()
        self.rows, self.cols, self.res = g["rows"], g["cols"], g["resolution"]
        self.pos, self.start = [0.0, 0.0], [0, 0]
        self._layers, self._order = {}, []
        for n in ("elevation", "elevation_min", "elevation_max"):  # ElevationMap's own layers
            self.add(n)

    def geometry(self):
        return types.SimpleNamespace(rows=self.rows, cols=self.cols, resolution=self.res,
                                     position=self.pos, start_index=self.start)

    def getSize(self):
        return (self.rows, self.cols)

    def rowStripe(self):
        return (0, self.rows)

    def getFrameId(self):
        return self._frame

    def setFrameId(self, f):
        self._frame = f

    def setPosition(self, p):
        self.pos = [float(p[0]), float(p[1])]

    def setStartIndex(self, s):
        self.start = [int(s[0]), int(s[1])]

    def getLayers(self):
        return list(self._order)

    def exists(self, n):
        return n in self._layers

    def add(self, n, fill=float("nan")):
        if n not in self._layers:
            self._order.append(n)
        self._layers[n] = np.full((self.rows, self.cols), fill, np.float32, order="F")

    def get(self, n):
        return self._layers[n].copy(order="F")

    def set(self, n, a):
        self._layers[n] = np.asfortranarray(np.asarray(a, np.float32))


def _fill_pair(make_map, width, height, res, frame, layers, pos=(0.0, 0.0), start=(0, 0), seed=0):
    """Same content in a product-side map and an oracle map."""
    rng = np.random.default_rng(seed)
    m = make_map(width, height, res, frame)
    o = ob.OracleMap(width, height, res)
    m.setPosition(pos)
    o.setPosition(pos)
    m.setStartIndex(start)
    o.setStartIndex(start)
    rows, cols = m.getSize()
    for n in layers:
        a = rng.standard_normal((rows, cols)).astype(np.float32)
        a[rng.random((rows, cols)) < 0.3] = np.nan
        if not m.exists(n):
            m.add(n)
        if not o.exists(n):
            o.add(n)
        m.set(n, a)
        o.set(n, a)
    return m, o


def _host(width, height, res, frame):
    return HostMap(width, height, res, frame)


CASES = [
    dict(width=5.0, height=5.0, res=0.5, frame="map", layers=["elevation"]),
    dict(width=10.0, height=10.0, res=1.0, frame="odom", layers=["elevation"], pos=(5.0, 3.0),
         start=(3, 7)),
    dict(width=4.0, height=6.0, res=0.1, frame='we"ird\\frame', layers=["elevation", "variance",
         "n_points", "color"], pos=(-12.3456789, 1e6 + 0.25), start=(11, 59)),
    dict(width=12.8, height=3.2, res=0.05, frame="", layers=["elevation", "state_upper_bound"],
         pos=(1e-7, -0.000123456789)),
]


@pytest.mark.parametrize("case", CASES)
def test_writer_bytes_equal_oracle(tmp_path, case):
    kw = {k: v for k, v in case.items() if k in ("pos", "start")}
    m, o = _fill_pair(_host, case["width"], case["height"], case["res"], case["frame"],
                      case["layers"], **kw)
    pa, pb = tmp_path / "a.npz", tmp_path / "b.npz"
    assert io_npz.saveNpz(str(pa), m)
    assert ob.save_npz(o, pb, case["frame"])
    assert pa.read_bytes() == pb.read_bytes()


@pytest.mark.parametrize("case", CASES)
def test_numpy_reads_it(tmp_path, case):
    kw = {k: v for k, v in case.items() if k in ("pos", "start")}
    m, _ = _fill_pair(_host, case["width"], case["height"], case["res"], case["frame"],
                      case["layers"], **kw)
    p = tmp_path / "a.npz"
    assert io_npz.saveNpz(str(p), m)
    with np.load(p) as z:
        assert sorted(z.files) == sorted(set(case["layers"]) | set(m.getLayers()) | {"meta"})
        for n in case["layers"]:
            np.testing.assert_array_equal(z[n], m.get(n))
            assert z[n].flags.f_contiguous
        assert z["meta"].shape == () and z["meta"].dtype.kind == "S"
    import zipfile
    with zipfile.ZipFile(p) as zf:
        assert zf.testzip() is None  # CRCs hold
        assert all(i.compress_type == zipfile.ZIP_STORED for i in zf.infolist())


@pytest.mark.parametrize("case", CASES)
def test_cross_load(tmp_path, case):
    kw = {k: v for k, v in case.items() if k in ("pos", "start")}
    m, o = _fill_pair(_host, case["width"], case["height"], case["res"], case["frame"],
                      case["layers"], **kw)
    pa, pb = tmp_path / "a.npz", tmp_path / "b.npz"
    assert io_npz.saveNpz(str(pa), m) and ob.save_npz(o, pb, case["frame"])
    # oracle reads the product's file, product reads the oracle's file; both must agree
    o2 = ob.OracleMap(1.0, 1.0, 0.5)
    ok, frame_o = ob.load_npz(o2, pa)
    assert ok
    m2 = HostMap()
    assert io_npz.loadNpz(str(pb), m2)
    g_o, g_m = o2.geometry(), m2.geometry()
    assert (g_o["rows"], g_o["cols"]) == (g_m.rows, g_m.cols) == m.getSize()
    assert g_o["resolution"] == g_m.resolution
    assert tuple(g_o["position"]) == tuple(g_m.position)
    assert tuple(g_o["start_index"]) == tuple(g_m.start_index)
    assert frame_o == m2.getFrameId()
    assert sorted(o2.layers()) == sorted(m2.getLayers())
    for n in case["layers"]:
        np.testing.assert_array_equal(o2.get(n), m.get(n))
        np.testing.assert_array_equal(m2.get(n), m.get(n))


def _roundtrip_suite(make_map, tmp_path):
    # RoundTrip (test_map_io.cpp)
    m = make_map(5.0, 5.0, 0.5, "map")
    m.setPosition((0.0, 0.0))
    e = np.full(m.getSize(), np.nan, np.float32)
    e[2, 3], e[0, 0], e[9, 9] = 1.5, -0.3, 42.0
    m.set("elevation", e)
    p = str(tmp_path / "rt.npz")
    assert io_npz.saveNpz(p, m, ["elevation"])
    l = make_map(1.0, 1.0, 0.5, "")
    assert io_npz.loadNpz(p, l)
    assert l.getSize() == (10, 10)
    assert l.geometry().resolution == pytest.approx(0.5)
    assert l.getFrameId() == "map"
    assert tuple(l.geometry().position) == (0.0, 0.0)
    np.testing.assert_array_equal(l.get("elevation"), e)

    # RoundTripWithStartIndex
    m = make_map(10.0, 10.0, 1.0, "odom")
    m.setPosition((5.0, 3.0))
    m.setStartIndex((3, 7))
    e = np.full(m.getSize(), np.nan, np.float32)
    e[0, 0], e[5, 5] = 10.0, 20.0
    m.set("elevation", e)
    p = str(tmp_path / "si.npz")
    assert io_npz.saveNpz(p, m, ["elevation"])
    l = make_map(1.0, 1.0, 0.5, "")
    assert io_npz.loadNpz(p, l)
    assert tuple(l.geometry().start_index) == (3, 7)
    assert tuple(l.geometry().position) == (5.0, 3.0)
    np.testing.assert_array_equal(l.get("elevation"), e)

    # MultipleLayers / SelectiveSave / NonexistentLayerSkipped
    m = make_map(5.0, 5.0, 0.5, "map")
    m.add("variance", 0.1)
    m.add("extra", 7.0)
    m.set("elevation", np.ones(m.getSize(), np.float32))
    p = str(tmp_path / "ml.npz")
    assert io_npz.saveNpz(p, m)
    l = make_map(1.0, 1.0, 0.5, "")
    assert io_npz.loadNpz(p, l)
    assert l.exists("variance") and l.exists("extra")
    assert l.get("variance")[0, 0] == np.float32(0.1)
    assert io_npz.saveNpz(p, m, ["elevation", "variance"])
    l = make_map(1.0, 1.0, 0.5, "")
    assert io_npz.loadNpz(p, l)
    assert l.exists("variance") and not l.exists("extra")
    assert io_npz.saveNpz(p, m, ["elevation", "no_such_layer"])
    l = make_map(1.0, 1.0, 0.5, "")
    assert io_npz.loadNpz(p, l)
    assert not l.exists("no_such_layer")

    # EmptyMap: all-NaN elevation still round-trips
    m = make_map(5.0, 5.0, 0.5, "map")
    p = str(tmp_path / "empty.npz")
    assert io_npz.saveNpz(p, m, ["elevation"])
    l = make_map(1.0, 1.0, 0.5, "")
    assert io_npz.loadNpz(p, l)
    assert np.isnan(l.get("elevation")).all()

    # metadata only → load reports failure (0 layers)
    assert io_npz.saveNpz(p, m, [])
    assert not io_npz.loadNpz(p, make_map(1.0, 1.0, 0.5, ""))

    # FutureVersionRejected: patch the version in place (same length, CRC not checked on load)
    p = tmp_path / "fv.npz"
    assert io_npz.saveNpz(str(p), m, ["elevation"])
    b = p.read_bytes()
    assert b'"version": 1' in b
    p.write_bytes(b.replace(b'"version": 1', b'"version":99'))
    assert not io_npz.loadNpz(str(p), make_map(1.0, 1.0, 0.5, ""))
    o = ob.OracleMap(1.0, 1.0, 0.5)
    assert not ob.load_npz(o, p)[0]

    # LoadNonExistentFile / SaveToInvalidPath / garbage
    assert not io_npz.loadNpz(str(tmp_path / "nope.npz"), make_map(1.0, 1.0, 0.5, ""))
    assert not io_npz.saveNpz(str(tmp_path / "no_dir" / "x.npz"), m)
    g = tmp_path / "garbage.npz"
    g.write_bytes(b"not a zip at all" * 10)
    assert not io_npz.loadNpz(str(g), make_map(1.0, 1.0, 0.5, ""))
    t = tmp_path / "trunc.npz"
    t.write_bytes(b[:200])
    assert not io_npz.loadNpz(str(t), make_map(1.0, 1.0, 0.5, ""))
    assert not ob.load_npz(o, t)[0]


def test_reference_cases_host(tmp_path):
    _roundtrip_suite(_host, tmp_path)


def test_shape_mismatch_entry_is_skipped(tmp_path):
    """A '<f4' entry whose shape differs from meta's size is ignored (io_npz.cpp:592-596)."""
    m = HostMap(5.0, 5.0, 0.5, "map")
    p = tmp_path / "a.npz"
    assert io_npz.saveNpz(str(p), m, ["elevation"])
    b = p.read_bytes()
    b2 = b.replace(b"'shape': (10, 10), }", b"'shape': (10, 20), }")
    p.write_bytes(b2)
    assert not io_npz.loadNpz(str(p), HostMap())  # only layer skipped → nothing loaded
    assert not ob.load_npz(ob.OracleMap(1.0, 1.0, 0.5), p)[0]


# ───────────────────────── device map ─────────────────────────

def _dev(width, height, res, frame):
    from fastdem_b200.api import ElevationMap
    return ElevationMap(width, height, res, frame)


@pytest.mark.gpu
def test_reference_cases_device(tmp_path):
    _roundtrip_suite(_dev, tmp_path)


@pytest.mark.gpu
def test_device_checkpoint_after_mapping(tmp_path):
    """Map a few scans (circular buffer rolled), checkpoint, restore into a fresh map, keep
    mapping on both: the restored map must continue bit-identically; the file must equal the
    oracle's for the same scans."""
    import fastdem_b200 as fdem
    from fastdem_b200 import synthetic as syn
    from parity_utils import run_pair

    wl = syn.WORKLOADS["tiny"]
    gmap, omap, gdem, odem, _, _ = run_pair(fdem, wl, 6)
    pa, pb = tmp_path / "g.npz", tmp_path / "o.npz"
    assert io_npz.saveNpz(str(pa), gmap)
    assert ob.save_npz(omap, pb, gmap.getFrameId())
    assert pa.read_bytes() == pb.read_bytes()

    restored = fdem.ElevationMap()
    assert io_npz.loadNpz(str(pa), restored)
    gdem2 = fdem.FastDEM(restored, wl.config())
    for k in range(6, 10):
        s = syn.make_scan(wl, k)
        cloud = fdem.PointCloud(s["xyzw"], s["intensity"], s["rgb"])
        a = gdem.integrate_stats(cloud, s["T_base_sensor"], s["T_world_base"])
        b = gdem2.integrate_stats(cloud, s["T_base_sensor"], s["T_world_base"])
        assert (a.n_kept, a.n_cells) == (b.n_kept, b.n_cells)
    assert sorted(restored.getLayers()) == sorted(gmap.getLayers())
    for n in gmap.getLayers():
        np.testing.assert_array_equal(restored.get(n), gmap.get(n))
