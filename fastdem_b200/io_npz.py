"""Map checkpoint IO — the reference's `fastdem::io::saveNpz / loadNpz`
(fastdem/include/fastdem/io/npz.hpp, fastdem/src/io_npz.cpp:376-612) for device-resident maps.

File format (byte-compatible with the reference; a file written here loads in the reference and
in `numpy.load`, and vice versa):

  * an uncompressed (STORE) ZIP archive: local header (sig 0x04034b50, version 20, all of
    flags/method/time/date zero, CRC-32, sizes, name), data, ...; central directory
    (sig 0x02014b50, version made-by/needed 20); end record (sig 0x06054b50);
  * one `<layer>.npy` per layer: NPY v1.0, dict `{'descr': '<f4', 'fortran_order': True,
    'shape': (rows, cols), }` space-padded so the data starts on a 64-byte boundary; the data is
    the layer's column-major buffer *as stored* (circular-buffer order, not unrolled);
  * `meta.npy`: a 0-d `|S<n>` array holding the JSON text
    `{"version": 1, "resolution": R, "position": [x, y], "frame_id": "...", "size": [r, c],
    "start_index": [a, b]}` with numbers in default-iostream (`%g`) formatting.

Layers come straight off the device (`fdem_map_layer_download`, one D2H copy per layer) and go
back with one H2D copy per layer; nothing here touches the oracle.
"""
from __future__ import annotations

import struct
import zlib
from typing import Optional, Sequence

import numpy as np

FORMAT_VERSION = 1  # io_npz.cpp:25
_MAX_ENTRIES = 1000  # io_npz.cpp:459
_MAX_NAME = 4096  # io_npz.cpp:474
_MAX_ENTRY_BYTES = 400_000_000  # io_npz.cpp:479


def _npy(dict_text: str, payload: bytes) -> bytes:
    pad = 64 - ((10 + len(dict_text) + 1) % 64)
    if pad == 64:
        pad = 0
    header = dict_text.encode("latin-1") + b" " * pad + b"\n"
    return b"\x93NUMPY\x01\x00" + struct.pack("<H", len(header)) + header + payload


def _npy_layer(a: np.ndarray) -> bytes:
    rows, cols = a.shape
    d = "{'descr': '<f4', 'fortran_order': True, 'shape': (%d, %d), }" % (rows, cols)
    return _npy(d, np.asfortranarray(a, dtype="<f4").tobytes(order="F"))


def _npy_string(s: bytes) -> bytes:
    return _npy("{'descr': '|S%d', 'fortran_order': False, 'shape': (), }" % len(s), s)


def _g(v: float) -> str:
    return "%g" % v  # == `ostream << v` at the default precision (6 significant digits)


def _meta_json(map) -> str:
    g = map.geometry()
    frame = map.getFrameId().replace("\\", "\\\\").replace('"', '\\"')
    return ('{"version": %d, "resolution": %s, "position": [%s, %s], "frame_id": "%s", '
            '"size": [%d, %d], "start_index": [%d, %d]}'
            % (FORMAT_VERSION, _g(g.resolution), _g(g.position[0]), _g(g.position[1]), frame,
               g.rows, g.cols, g.start_index[0], g.start_index[1]))


def saveNpz(filename: str, map, layer_names: Optional[Sequence[str]] = None) -> bool:
    """io_npz.cpp:376-437.  Returns False (never raises) when the file cannot be written;
    names that are not layers of `map` are skipped; an empty list saves only the metadata."""
    if map.rowStripe() != (0, map.getSize()[0]):
        raise ValueError("saveNpz needs the whole map; gather the stripes first")
    names = list(map.getLayers()) if layer_names is None else list(layer_names)
    try:
        fs = open(filename, "wb")
    except OSError:
        return False
    entries = []  # (name, crc, size, offset)
    try:
        with fs:
            def add(name: str, blob: bytes) -> None:
                nm = name.encode()
                crc = zlib.crc32(blob) & 0xFFFFFFFF
                entries.append((nm, crc, len(blob), fs.tell()))
                fs.write(struct.pack("<IHHHHHIIIHH", 0x04034B50, 20, 0, 0, 0, 0, crc, len(blob),
                                     len(blob), len(nm), 0))
                fs.write(nm)
                fs.write(blob)

            for n in names:
                if not map.exists(n):
                    continue
                add(n + ".npy", _npy_layer(map.get(n)))
            add("meta.npy", _npy_string(_meta_json(map).encode()))
            cd_offset = fs.tell()
            for nm, crc, size, off in entries:
                fs.write(struct.pack("<IHHHHHHIIIHHHHHII", 0x02014B50, 20, 20, 0, 0, 0, 0, crc,
                                     size, size, len(nm), 0, 0, 0, 0, 0, off))
                fs.write(nm)
            cd_size = fs.tell() - cd_offset
            fs.write(struct.pack("<IHHHHIIH", 0x06054B50, 0, 0, len(entries), len(entries),
                                 cd_size, cd_offset, 0))
    except OSError:
        return False
    return True


def _stof(s: str) -> float:
    """std::stof: leading whitespace, then the longest numeric prefix; raises if none."""
    import re
    m = re.match(r"\s*[+-]?(?:\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?|inf(?:inity)?|nan)",
                 s, re.IGNORECASE)
    if not m:
        raise ValueError(s)
    return float(np.float32(float(m.group(0))))


def _stoi(s: str) -> int:
    import re
    m = re.match(r"\s*[+-]?\d+", s)
    if not m:
        raise ValueError(s)
    return int(m.group(0))


def _json_float(j: str, key: str):
    pos = j.find('"%s"' % key)
    if pos < 0:
        return None
    pos = j.find(":", pos)
    if pos < 0:
        return None
    try:
        return _stof(j[pos + 1:])
    except ValueError:
        return None


def _json_pair(j: str, key: str, conv):
    pos = j.find('"%s"' % key)
    if pos < 0:
        return None
    pos = j.find("[", pos)
    if pos < 0:
        return None
    end = j.find("]", pos)
    if end < 0:
        return None
    inner = j[pos + 1:end]
    comma = inner.find(",")
    if comma < 0:
        return None
    try:
        return conv(inner[:comma]), conv(inner[comma + 1:])
    except ValueError:
        return None


def _json_string(j: str, key: str):
    pos = j.find('"%s"' % key)
    if pos < 0:
        return None
    pos = j.find(":", pos)
    if pos < 0:
        return None
    q1 = j.find('"', pos + 1)
    if q1 < 0:
        return None
    q2 = j.find('"', q1 + 1)
    if q2 < 0:
        return None
    return j[q1 + 1:q2]


def _parse_npy(buf: bytes):
    """-> dict(kind='f4', rows, cols, off) | dict(kind='S', slen, off) | None (io_npz.cpp:310-361)."""
    if len(buf) < 10 or buf[:6] != b"\x93NUMPY":
        return None
    (hl,) = struct.unpack_from("<H", buf, 8)
    off = 10 + hl
    if off > len(buf):
        return None
    d = buf[10:off].decode("latin-1")
    if "'descr'" not in d:
        return None
    try:
        if "'<f4'" in d:
            sp = d.find("'shape'")
            p0 = d.find("(", sp) if sp >= 0 else -1
            p1 = d.find(")", p0) if p0 >= 0 else -1
            if p1 < 0:
                return None
            shape = d[p0 + 1:p1]
            comma = shape.find(",")
            if comma < 0:
                return None
            c = shape[comma + 1:]
            if c.endswith(","):
                c = c[:-1]
            c = c.lstrip(" ")
            if not c:
                return None
            return dict(kind="f4", rows=_stoi(shape[:comma]), cols=_stoi(c), off=off)
        if "'|S" in d:
            s0 = d.find("'|S")
            s1 = d.find("'", s0 + 3)
            if s1 < 0:
                return None
            return dict(kind="S", slen=_stoi(d[s0 + 3:s1]), off=off)
    except ValueError:
        return None
    return None


def loadNpz(filename: str, map) -> bool:
    """io_npz.cpp:440-612.  Re-creates `map`'s geometry (`setGeometry(res*rows, res*cols, res)`
    in float32, then position, start index, frame id) and uploads every '<f4' entry whose shape
    matches.  False when the file is missing/corrupt, has no meta.npy, a format version newer
    than 1, or no loadable layer."""
    try:
        with open(filename, "rb") as fs:
            blob = fs.read()
    except OSError:
        return False
    entries = []
    pos = 0
    while len(entries) < _MAX_ENTRIES:
        if pos + 30 > len(blob) or struct.unpack_from("<I", blob, pos)[0] != 0x04034B50:
            break
        usize, nlen, xlen = struct.unpack_from("<IHH", blob, pos + 22)
        if nlen > _MAX_NAME or usize > _MAX_ENTRY_BYTES:
            return False
        name = blob[pos + 30:pos + 30 + nlen]
        start = pos + 30 + nlen + xlen
        if len(name) < nlen or start + usize > len(blob):
            return False
        entries.append((name.decode("latin-1"), blob[start:start + usize]))
        pos = start + usize
    if not entries:
        return False

    meta = None
    for name, data in entries:
        if name != "meta.npy":
            continue
        info = _parse_npy(data)
        if info is None or info["kind"] != "S" or info["off"] + info["slen"] > len(data):
            return False
        meta = data[info["off"]:info["off"] + info["slen"]].decode("latin-1")
        break
    if meta is None:
        return False

    version = _json_float(meta, "version")
    if version is not None and int(version) > FORMAT_VERSION:
        return False
    resolution = _json_float(meta, "resolution")
    position = _json_pair(meta, "position", _stof)
    size = _json_pair(meta, "size", _stoi)
    if resolution is None or position is None or size is None:
        return False
    frame = _json_string(meta, "frame_id")
    start_index = _json_pair(meta, "start_index", _stoi) or (0, 0)
    rows, cols = size
    if rows <= 0 or cols <= 0 or resolution <= 0:
        return False

    res32 = np.float32(resolution)
    map.setGeometry(float(res32 * np.float32(rows)), float(res32 * np.float32(cols)), float(res32))
    map.setPosition(position)
    map.setStartIndex(start_index)
    map.setFrameId(frame or "")

    loaded = 0
    expected = rows * cols * 4
    for name, data in entries:
        if name == "meta.npy" or len(name) <= 4 or not name.endswith(".npy"):
            continue
        info = _parse_npy(data)
        if info is None or info["kind"] != "f4":
            continue
        if info["rows"] != rows or info["cols"] != cols or info["off"] + expected > len(data):
            continue
        lname = name[:-4]
        if not map.exists(lname):
            map.add(lname)
        a = np.frombuffer(data, dtype="<f4", count=rows * cols, offset=info["off"])
        map.set(lname, a.reshape((rows, cols), order="F"))
        loaded += 1
    return loaded > 0
