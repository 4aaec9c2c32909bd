// io_npz.hpp — fastdem::io::saveNpz / loadNpz for C++ callers
// (fastdem/include/fastdem/io/npz.hpp, fastdem/src/io_npz.cpp:141-238, 376-612).
//
// Same file format as the reference's writer (and fastdem_b200/io_npz.py): an uncompressed ZIP
// (STORE) with one NPY v1.0 entry per layer — '<f4', fortran_order True, shape (rows, cols), the
// column-major buffer exactly as stored (circular-buffer order, not unrolled), header padded to a
// 64-byte boundary — plus meta.npy, a 0-d '|S<n>' array holding the JSON text
//   {"version": 1, "resolution": R, "position": [x, y], "frame_id": "...", "size": [r, c],
//    "start_index": [a, b]}
// with numbers in default-iostream formatting.  Layers come off the device with one D2H copy each
// (ElevationMap::get) and go back with one H2D copy each.  loadNpz calls setGeometry on the map
// it restores into — in place, so a FastDEM already bound to that map keeps working.
#pragma once

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "fastdem/elevation_map.hpp"

namespace fastdem {
namespace io {
namespace detail {

constexpr int kFormatVersion = 1;            // io_npz.cpp:25
constexpr size_t kMaxEntries = 1000;         // io_npz.cpp:459
constexpr size_t kMaxName = 4096;            // io_npz.cpp:474
constexpr size_t kMaxEntryBytes = 400000000; // io_npz.cpp:479

inline uint32_t crc32(const uint8_t* p, size_t n) {
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : (c >> 1);
      table[i] = c;
    }
    init = true;
  }
  uint32_t c = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

inline void put16(std::string& s, uint16_t v) { s.push_back(static_cast<char>(v & 0xFF)); s.push_back(static_cast<char>(v >> 8)); }
inline void put32(std::string& s, uint32_t v) { for (int i = 0; i < 4; ++i) s.push_back(static_cast<char>((v >> (8 * i)) & 0xFF)); }
inline uint16_t get16(const uint8_t* p) { return static_cast<uint16_t>(p[0] | (p[1] << 8)); }
inline uint32_t get32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | (static_cast<uint32_t>(p[3]) << 24); }

inline std::string npy(const std::string& dict, const void* payload, size_t bytes) {
  size_t pad = 64 - ((10 + dict.size() + 1) % 64);
  if (pad == 64) pad = 0;
  std::string header = dict + std::string(pad, ' ') + "\n";
  std::string out("\x93NUMPY\x01\x00", 8);
  put16(out, static_cast<uint16_t>(header.size()));
  out += header;
  out.append(static_cast<const char*>(payload), bytes);
  return out;
}

inline std::string num(double v) {  // `ostream << v` at the default precision
  std::ostringstream o;
  o << v;
  return o.str();
}

struct NpyInfo {
  char kind = 0;  // 'f' = '<f4' matrix, 'S' = byte string
  int rows = 0, cols = 0;
  size_t slen = 0, off = 0;
};

inline bool parseNpy(const uint8_t* buf, size_t n, NpyInfo& info) {  // io_npz.cpp:310-361
  if (n < 10 || std::memcmp(buf, "\x93NUMPY", 6) != 0) return false;
  const size_t off = 10 + get16(buf + 8);
  if (off > n) return false;
  const std::string d(reinterpret_cast<const char*>(buf + 10), off - 10);
  if (d.find("'descr'") == std::string::npos) return false;
  info.off = off;
  if (d.find("'<f4'") != std::string::npos) {
    const size_t sp = d.find("'shape'");
    const size_t p0 = sp == std::string::npos ? sp : d.find('(', sp);
    const size_t p1 = p0 == std::string::npos ? p0 : d.find(')', p0);
    if (p1 == std::string::npos) return false;
    const std::string shape = d.substr(p0 + 1, p1 - p0 - 1);
    const size_t comma = shape.find(',');
    if (comma == std::string::npos) return false;
    char* e = nullptr;
    const long r = std::strtol(shape.c_str(), &e, 10);
    if (e == shape.c_str()) return false;
    const char* cs = shape.c_str() + comma + 1;
    const long c = std::strtol(cs, &e, 10);
    if (e == cs) return false;
    info.kind = 'f';
    info.rows = static_cast<int>(r);
    info.cols = static_cast<int>(c);
    return true;
  }
  const size_t s0 = d.find("'|S");
  if (s0 != std::string::npos) {
    char* e = nullptr;
    const long l = std::strtol(d.c_str() + s0 + 3, &e, 10);
    if (e == d.c_str() + s0 + 3 || l < 0) return false;
    info.kind = 'S';
    info.slen = static_cast<size_t>(l);
    return true;
  }
  return false;
}

inline bool jsonNumber(const std::string& j, const std::string& key, double& v) {
  size_t pos = j.find("\"" + key + "\"");
  if (pos == std::string::npos) return false;
  pos = j.find(':', pos);
  if (pos == std::string::npos) return false;
  char* e = nullptr;
  v = std::strtod(j.c_str() + pos + 1, &e);
  return e != j.c_str() + pos + 1;
}
inline bool jsonPair(const std::string& j, const std::string& key, double& a, double& b) {
  size_t pos = j.find("\"" + key + "\"");
  if (pos == std::string::npos) return false;
  pos = j.find('[', pos);
  const size_t end = pos == std::string::npos ? pos : j.find(']', pos);
  if (end == std::string::npos) return false;
  const std::string inner = j.substr(pos + 1, end - pos - 1);
  const size_t comma = inner.find(',');
  if (comma == std::string::npos) return false;
  char* e = nullptr;
  a = std::strtod(inner.c_str(), &e);
  if (e == inner.c_str()) return false;
  const char* s2 = inner.c_str() + comma + 1;
  b = std::strtod(s2, &e);
  return e != s2;
}
inline bool jsonString(const std::string& j, const std::string& key, std::string& out) {
  size_t pos = j.find("\"" + key + "\"");
  if (pos == std::string::npos) return false;
  pos = j.find(':', pos);
  if (pos == std::string::npos) return false;
  const size_t q1 = j.find('"', pos + 1);
  const size_t q2 = q1 == std::string::npos ? q1 : j.find('"', q1 + 1);
  if (q2 == std::string::npos) return false;
  out = j.substr(q1 + 1, q2 - q1 - 1);
  return true;
}

}  // namespace detail

// io_npz.cpp:376-437.  false (never throws) when the file cannot be written; names that are not
// layers of the map are skipped; an empty list saves only the metadata; no list = every layer.
inline bool saveNpz(const std::string& filename, const ElevationMap& map,
                    const std::vector<std::string>* layer_names = nullptr) {
  using namespace detail;
  const std::vector<std::string> names = layer_names ? *layer_names : map.getLayers();
  std::ofstream fs(filename, std::ios::binary);
  if (!fs) return false;
  struct Entry { std::string name; uint32_t crc, size, offset; };
  std::vector<Entry> entries;
  uint32_t offset = 0;
  auto add = [&](const std::string& name, const std::string& blob) {
    const uint32_t crc = crc32(reinterpret_cast<const uint8_t*>(blob.data()), blob.size());
    entries.push_back({name, crc, static_cast<uint32_t>(blob.size()), offset});
    std::string h;
    put32(h, 0x04034B50u); put16(h, 20); put16(h, 0); put16(h, 0); put16(h, 0); put16(h, 0);
    put32(h, crc); put32(h, static_cast<uint32_t>(blob.size())); put32(h, static_cast<uint32_t>(blob.size()));
    put16(h, static_cast<uint16_t>(name.size())); put16(h, 0);
    fs.write(h.data(), static_cast<std::streamsize>(h.size()));
    fs.write(name.data(), static_cast<std::streamsize>(name.size()));
    fs.write(blob.data(), static_cast<std::streamsize>(blob.size()));
    offset += static_cast<uint32_t>(h.size() + name.size() + blob.size());
  };
  const nanogrid::Size size = map.getSize();
  for (const auto& n : names) {
    if (!map.exists(n)) continue;
    const nanogrid::Matrix m = map.get(n);
    const std::string dict = "{'descr': '<f4', 'fortran_order': True, 'shape': (" + std::to_string(m.rows()) +
                             ", " + std::to_string(m.cols()) + "), }";
    add(n + ".npy", npy(dict, m.data(), m.size() * sizeof(float)));
  }
  {
    std::string frame;
    for (char ch : map.getFrameId()) {
      if (ch == '\\' || ch == '"') frame.push_back('\\');
      frame.push_back(ch);
    }
    const nanogrid::Position p = map.getPosition();
    const nanogrid::Index st = map.getStartIndex();
    const std::string meta = "{\"version\": " + std::to_string(kFormatVersion) + ", \"resolution\": " +
                             num(map.getResolution()) + ", \"position\": [" + num(p(0)) + ", " + num(p(1)) +
                             "], \"frame_id\": \"" + frame + "\", \"size\": [" + std::to_string(size(0)) + ", " +
                             std::to_string(size(1)) + "], \"start_index\": [" + std::to_string(st(0)) + ", " +
                             std::to_string(st(1)) + "]}";
    const std::string dict = "{'descr': '|S" + std::to_string(meta.size()) + "', 'fortran_order': False, 'shape': (), }";
    add("meta.npy", npy(dict, meta.data(), meta.size()));
  }
  const uint32_t cd_offset = offset;
  uint32_t cd_size = 0;
  for (const auto& e : entries) {
    std::string h;
    put32(h, 0x02014B50u); put16(h, 20); put16(h, 20); put16(h, 0); put16(h, 0); put16(h, 0); put16(h, 0);
    put32(h, e.crc); put32(h, e.size); put32(h, e.size);
    put16(h, static_cast<uint16_t>(e.name.size())); put16(h, 0); put16(h, 0); put16(h, 0); put16(h, 0);
    put32(h, 0); put32(h, e.offset);
    fs.write(h.data(), static_cast<std::streamsize>(h.size()));
    fs.write(e.name.data(), static_cast<std::streamsize>(e.name.size()));
    cd_size += static_cast<uint32_t>(h.size() + e.name.size());
  }
  std::string endrec;
  put32(endrec, 0x06054B50u); put16(endrec, 0); put16(endrec, 0);
  put16(endrec, static_cast<uint16_t>(entries.size())); put16(endrec, static_cast<uint16_t>(entries.size()));
  put32(endrec, cd_size); put32(endrec, cd_offset); put16(endrec, 0);
  fs.write(endrec.data(), static_cast<std::streamsize>(endrec.size()));
  return static_cast<bool>(fs);
}

// io_npz.cpp:440-612.  Re-creates the map's geometry (setGeometry(res*rows, res*cols, res) in
// float32, then position, start index, frame id) and uploads every '<f4' entry whose shape
// matches.  false when the file is missing / corrupt, has no meta.npy, a format version newer
// than 1, or no loadable layer.
inline bool loadNpz(const std::string& filename, ElevationMap& map) {
  using namespace detail;
  std::ifstream fs(filename, std::ios::binary);
  if (!fs) return false;
  const std::string blob((std::istreambuf_iterator<char>(fs)), std::istreambuf_iterator<char>());
  const uint8_t* b = reinterpret_cast<const uint8_t*>(blob.data());
  struct Entry { std::string name; size_t start, size; };
  std::vector<Entry> entries;
  size_t pos = 0;
  while (entries.size() < kMaxEntries) {
    if (pos + 30 > blob.size() || get32(b + pos) != 0x04034B50u) break;
    const size_t usize = get32(b + pos + 22);
    const size_t nlen = get16(b + pos + 26), xlen = get16(b + pos + 28);
    if (nlen > kMaxName || usize > kMaxEntryBytes) return false;
    const size_t start = pos + 30 + nlen + xlen;
    if (pos + 30 + nlen > blob.size() || start + usize > blob.size()) return false;
    entries.push_back({blob.substr(pos + 30, nlen), start, usize});
    pos = start + usize;
  }
  if (entries.empty()) return false;
  std::string meta;
  bool have_meta = false;
  for (const auto& e : entries) {
    if (e.name != "meta.npy") continue;
    NpyInfo info;
    if (!parseNpy(b + e.start, e.size, info) || info.kind != 'S' || info.off + info.slen > e.size) return false;
    meta.assign(reinterpret_cast<const char*>(b + e.start + info.off), info.slen);
    have_meta = true;
    break;
  }
  if (!have_meta) return false;
  double version = 0.0, res = 0.0, px = 0.0, py = 0.0, sr = 0.0, sc = 0.0, ir = 0.0, ic = 0.0;
  if (jsonNumber(meta, "version", version) && static_cast<int>(version) > kFormatVersion) return false;
  if (!jsonNumber(meta, "resolution", res) || !jsonPair(meta, "size", sr, sc)) return false;
  const float resolution = static_cast<float>(res);
  const int rows = static_cast<int>(sr), cols = static_cast<int>(sc);
  if (!(resolution > 0.0f) || rows <= 0 || cols <= 0) return false;
  map.setGeometry(resolution * static_cast<float>(rows), resolution * static_cast<float>(cols), resolution);
  if (jsonPair(meta, "position", px, py)) map.setPosition(nanogrid::Position(px, py));
  if (jsonPair(meta, "start_index", ir, ic)) map.setStartIndex(nanogrid::Index(static_cast<int>(ir), static_cast<int>(ic)));
  std::string frame;
  if (jsonString(meta, "frame_id", frame)) map.setFrameId(frame);
  int loaded = 0;
  for (const auto& e : entries) {
    if (e.name == "meta.npy" || e.name.size() < 5 || e.name.compare(e.name.size() - 4, 4, ".npy") != 0) continue;
    NpyInfo info;
    if (!parseNpy(b + e.start, e.size, info) || info.kind != 'f') continue;
    if (info.rows != rows || info.cols != cols) continue;
    const size_t bytes = static_cast<size_t>(rows) * cols * sizeof(float);
    if (info.off + bytes > e.size) continue;
    nanogrid::Matrix m(rows, cols);
    std::memcpy(m.data(), b + e.start + info.off, bytes);
    const std::string layer_name = e.name.substr(0, e.name.size() - 4);
    if (!map.exists(layer_name)) map.add(layer_name);
    map.set(layer_name, m);
    ++loaded;
  }
  return loaded > 0;
}

}  // namespace io
}  // namespace fastdem
