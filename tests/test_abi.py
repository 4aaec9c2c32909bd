"""CPU-side checks of the drop-in boundary: libfastdem_b200.so loads without a GPU, exports
every symbol include/fastdem_b200.h declares, struct layouts agree between the header, the
ctypes mirror and the oracle's mirror, and the host-side geometry code (grid_geom.h) agrees
with the oracle's independent restatement.  No compute calls (there is no GPU here)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle_binding as ob
from fastdem_b200 import capi

REPO = Path(__file__).resolve().parent.parent
HEADER = (REPO / "include" / "fastdem_b200.h").read_text()


def declared_functions():
    names = re.findall(r"^(?:fdem_status|int32_t|const char\*|void\*|void)\s+(fdem_[a-z0-9_]+)\s*\(",
                       HEADER, flags=re.M)
    return sorted(set(names))


def test_library_loads_and_reports_abi_version():
    lib = capi.load_library()
    assert lib.fdem_abi_version() == 1
    assert lib.fdem_status_string(0) == b"ok"


def test_every_declared_symbol_is_exported_and_bound():
    declared = declared_functions()
    assert len(declared) >= 45
    out = subprocess.run(["nm", "-D", "--defined-only", str(capi.LIB_PATH)], check=True,
                         capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (fdem_[a-z0-9_]+)", out))
    missing = [n for n in declared if n not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    unbound = [n for n in declared if n not in capi.SIGNATURES]
    assert not unbound, f"declared in the header but not in capi.SIGNATURES: {unbound}"
    stale = [n for n in capi.SIGNATURES if n not in declared]
    assert not stale, f"bound in capi.py but not declared in the header: {stale}"
    # nothing else leaks out of the library (built with -fvisibility=hidden)
    assert not [n for n in exported if n not in declared]


def header_struct_fields(name):
    m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), HEADER, flags=re.S)
    assert m, name
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        typ, rest = decl.split(None, 1)
        for item in rest.split(","):
            item = item.strip()
            mm = re.match(r"([a-zA-Z0-9_]+)(?:\[(\d+)\])?$", item)
            assert mm, item
            fields.append((mm.group(1), typ, int(mm.group(2) or 1)))
    return fields


CT = {"float": C.c_float, "int32_t": C.c_int32, "int64_t": C.c_int64, "double": C.c_double}


@pytest.mark.parametrize("cname,cls", [("fdem_config", capi.FdemConfig),
                                       ("fdem_scan_stats", capi.FdemScanStats),
                                       ("fdem_geometry", capi.FdemGeometry)])
def test_struct_layout_matches_header(cname, cls):
    want = header_struct_fields(cname)
    got = cls._fields_
    assert [f[0] for f in got] == [w[0] for w in want]
    for (gname, gtype), (wname, wtype, wlen) in zip(got, want):
        base = CT[wtype]
        assert gtype is (base if wlen == 1 else base * wlen) or C.sizeof(gtype) == C.sizeof(base) * wlen


def test_default_config_matches_reference_defaults_and_oracle():
    a, b = capi.default_config(), ob.default_config()
    assert bytes(a) == bytes(b)
    # fastdem/include/fastdem/config/*.hpp
    assert a.z_min == -np.finfo(np.float32).max and a.z_max == np.finfo(np.float32).max
    assert a.range_min == 0.0 and a.range_max == np.finfo(np.float32).max
    assert a.sensor_type == capi.SENSOR_LIDAR
    assert (a.lidar_range_noise, a.lidar_angular_noise) == (np.float32(0.02), np.float32(0.001))
    assert a.mode == capi.MODE_LOCAL and a.estimation_type == capi.EST_KALMAN
    assert list(a.p2_dn) == [np.float32(v) for v in (0.01, 0.16, 0.5, 0.84, 0.99)]
    assert a.p2_elevation_marker == 3 and a.raycasting_enabled == 0
    assert a.rc_clear_threshold == -1.0 and a.rc_log_odds_max == 2.0


def test_calls_fail_loudly_without_a_gpu():
    """No CPU fallback: creating a map without a usable device is an error, not a detour."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device error path cannot be exercised")
    lib = capi.load_library()
    h = C.c_void_p()
    st = lib.fdem_map_create(10.0, 10.0, 0.5, 0, None, C.byref(h))
    assert st != capi.FDEM_OK and not h.value
    assert lib.fdem_last_error()
    with pytest.raises(capi.FdemError):
        capi.check(st)
    import fastdem_b200 as fd
    with pytest.raises(capi.FdemError):
        fd.ElevationMap(10.0, 10.0, 0.5)


def test_product_never_references_the_oracle():
    """oracle/ is test infrastructure: nothing under fastdem_b200/ or include/ may touch it."""
    offenders = []
    for p in list((REPO / "fastdem_b200").rglob("*")) + list((REPO / "include").rglob("*")):
        if p.is_file() and p.suffix in {".py", ".cu", ".cuh", ".h", ".hpp", ".inc", ".cpp"}:
            txt = p.read_text(errors="replace")
            if re.search(r"oracle_binding|libfdem_oracle|fdem_oracle\.hpp|[\"'/]oracle/", txt):
                if p.name == "_build.py":  # builds the checker; never loads it
                    continue
                offenders.append(str(p.relative_to(REPO)))
    assert not offenders, offenders
