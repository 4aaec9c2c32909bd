// fdem_oracle_io.cpp — CPU ORACLE (test infrastructure): restatement of the reference's map
// checkpoint format, fastdem/src/io_npz.cpp (saveNpz :376-437, loadNpz :440-612): every layer
// as a Fortran-order '<f4' .npy (v1.0 header padded to 64 bytes) inside an uncompressed
// (STORE) ZIP with CRC-32, plus meta.npy = a 0-d '|S<n>' string holding
// {"version", "resolution", "position", "frame_id", "size", "start_index"} as JSON written
// with default iostream float formatting.
#include <array>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "fdem_oracle.hpp"

using namespace fdem_oracle;

namespace {

uint32_t crc32(const void* data, size_t len) {  // io_npz.cpp:29-50
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int j = 0; j < 8; ++j) c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
      table[i] = c;
    }
    init = true;
  }
  const auto* buf = static_cast<const uint8_t*>(data);
  uint32_t crc = 0xFFFFFFFFu;
  for (size_t i = 0; i < len; ++i) crc = table[(crc ^ buf[i]) & 0xFF] ^ (crc >> 8);
  return crc ^ 0xFFFFFFFFu;
}
void w16(std::ostream& os, uint16_t v) { os.put(char(v & 0xFF)); os.put(char((v >> 8) & 0xFF)); }
void w32(std::ostream& os, uint32_t v) { for (int i = 0; i < 4; ++i) os.put(char((v >> (8 * i)) & 0xFF)); }
uint16_t r16(std::istream& is) { uint8_t b[2]; is.read(reinterpret_cast<char*>(b), 2); return uint16_t(b[0] | (b[1] << 8)); }
uint32_t r32(std::istream& is) {
  uint8_t b[4];
  is.read(reinterpret_cast<char*>(b), 4);
  return uint32_t(b[0]) | (uint32_t(b[1]) << 8) | (uint32_t(b[2]) << 16) | (uint32_t(b[3]) << 24);
}

struct ZipEntry { std::string name; uint32_t crc, size, offset; };

void localHeader(std::ostream& os, const ZipEntry& e) {  // :80-93
  w32(os, 0x04034b50); w16(os, 20); w16(os, 0); w16(os, 0); w16(os, 0); w16(os, 0);
  w32(os, e.crc); w32(os, e.size); w32(os, e.size);
  w16(os, uint16_t(e.name.size())); w16(os, 0);
  os.write(e.name.data(), e.name.size());
}
void centralHeader(std::ostream& os, const ZipEntry& e) {  // :95-114
  w32(os, 0x02014b50); w16(os, 20); w16(os, 20); w16(os, 0); w16(os, 0); w16(os, 0); w16(os, 0);
  w32(os, e.crc); w32(os, e.size); w32(os, e.size);
  w16(os, uint16_t(e.name.size())); w16(os, 0); w16(os, 0); w16(os, 0); w16(os, 0);
  w32(os, 0); w32(os, e.offset);
  os.write(e.name.data(), e.name.size());
}

std::vector<char> npyHeader(std::string dict) {  // shared tail of buildNpyArray / buildNpyString
  size_t padding = 64 - ((10 + dict.size() + 1) % 64);
  if (padding == 64) padding = 0;
  dict.append(padding, ' ');
  dict.push_back('\n');
  const uint16_t hl = uint16_t(dict.size());
  std::vector<char> buf;
  const char magic[] = {'\x93', 'N', 'U', 'M', 'P', 'Y', '\x01', '\x00'};
  buf.insert(buf.end(), magic, magic + 8);
  buf.push_back(char(hl & 0xFF));
  buf.push_back(char((hl >> 8) & 0xFF));
  buf.insert(buf.end(), dict.begin(), dict.end());
  return buf;
}
std::vector<char> npyArray(const float* data, int rows, int cols) {  // :139-170
  std::ostringstream d;
  d << "{'descr': '<f4', 'fortran_order': True, 'shape': (" << rows << ", " << cols << "), }";
  std::vector<char> buf = npyHeader(d.str());
  const char* raw = reinterpret_cast<const char*>(data);
  buf.insert(buf.end(), raw, raw + size_t(rows) * cols * 4);
  return buf;
}
std::vector<char> npyString(const std::string& s) {  // :172-197
  std::ostringstream d;
  d << "{'descr': '|S" << s.size() << "', 'fortran_order': False, 'shape': (), }";
  std::vector<char> buf = npyHeader(d.str());
  buf.insert(buf.end(), s.begin(), s.end());
  return buf;
}
std::string escapeJson(const std::string& s) {  // :199-211
  std::string out;
  for (char c : s) { if (c == '"') out += "\\\""; else if (c == '\\') out += "\\\\"; else out += c; }
  return out;
}
std::string metaJson(const ElevationMap& m, const std::string& frame) {  // :214-227
  std::ostringstream j;
  j << "{\"version\": " << 1 << ", \"resolution\": " << m.resolution() << ", \"position\": ["
    << m.position()[0] << ", " << m.position()[1] << "]" << ", \"frame_id\": \"" << escapeJson(frame)
    << "\"" << ", \"size\": [" << m.rows() << ", " << m.cols() << "]" << ", \"start_index\": ["
    << m.startIndex().r << ", " << m.startIndex().c << "]" << "}";
  return j.str();
}

bool jsonFloat(const std::string& j, const std::string& key, float& out) {  // :231-242
  auto pos = j.find("\"" + key + "\"");
  if (pos == std::string::npos) return false;
  pos = j.find(':', pos);
  if (pos == std::string::npos) return false;
  try { out = std::stof(j.substr(pos + 1)); } catch (...) { return false; }
  return true;
}
template <typename T, typename F>
bool jsonPair(const std::string& j, const std::string& key, T& a, T& b, F conv) {  // :244-283
  auto pos = j.find("\"" + key + "\"");
  if (pos == std::string::npos) return false;
  pos = j.find('[', pos);
  if (pos == std::string::npos) return false;
  auto end = j.find(']', pos);
  if (end == std::string::npos) return false;
  std::string inner = j.substr(pos + 1, end - pos - 1);
  auto comma = inner.find(',');
  if (comma == std::string::npos) return false;
  try { a = conv(inner.substr(0, comma)); b = conv(inner.substr(comma + 1)); } catch (...) { return false; }
  return true;
}
bool jsonString(const std::string& j, const std::string& key, std::string& out) {  // :285-298
  auto pos = j.find("\"" + key + "\"");
  if (pos == std::string::npos) return false;
  pos = j.find(':', pos);
  if (pos == std::string::npos) return false;
  auto q1 = j.find('"', pos + 1);
  if (q1 == std::string::npos) return false;
  auto q2 = j.find('"', q1 + 1);
  if (q2 == std::string::npos) return false;
  out = j.substr(q1 + 1, q2 - q1 - 1);
  return true;
}

struct NpyInfo { int rows = 0, cols = 0; bool is_float = false, is_string = false; size_t slen = 0, off = 0; };
bool parseNpy(const char* buf, size_t n, NpyInfo& info) {  // :310-361
  if (n < 10 || buf[0] != '\x93' || std::memcmp(buf + 1, "NUMPY", 5) != 0) return false;
  const uint16_t hl = uint16_t(uint8_t(buf[8]) | (uint8_t(buf[9]) << 8));
  info.off = 10 + hl;
  if (info.off > n) return false;
  std::string dict(buf + 10, hl);
  if (dict.find("'descr'") == std::string::npos) return false;
  if (dict.find("'<f4'") != std::string::npos) {
    info.is_float = true;
    auto sp = dict.find("'shape'");
    if (sp == std::string::npos) return false;
    auto p0 = dict.find('(', sp);
    if (p0 == std::string::npos) return false;
    auto p1 = dict.find(')', p0);
    if (p1 == std::string::npos) return false;
    std::string shape = dict.substr(p0 + 1, p1 - p0 - 1);
    auto comma = shape.find(',');
    if (comma == std::string::npos) return false;
    try {
      info.rows = std::stoi(shape.substr(0, comma));
      std::string c = shape.substr(comma + 1);
      if (!c.empty() && c.back() == ',') c.pop_back();
      while (!c.empty() && c.front() == ' ') c.erase(0, 1);
      if (c.empty()) return false;
      info.cols = std::stoi(c);
    } catch (...) { return false; }
  } else if (dict.find("'|S") != std::string::npos) {
    info.is_string = true;
    auto s0 = dict.find("'|S");
    auto s1 = dict.find('\'', s0 + 3);
    if (s1 == std::string::npos) return false;
    try { info.slen = std::stoul(dict.substr(s0 + 3, s1 - s0 - 3)); } catch (...) { return false; }
  } else {
    return false;
  }
  return true;
}

}  // namespace

extern "C" {

// saveNpz(filename, map, layer_names): layer_names = '\n'-separated list, or null for all
// layers in creation order.  Returns 1 / 0 like the reference's bool.
int orc_save_npz(void* mp, const char* filename, const char* frame_id, const char* layer_names) {
  auto& map = *static_cast<ElevationMap*>(mp);
  std::ofstream fs(filename, std::ios::binary);
  if (!fs.is_open()) return 0;
  std::vector<std::string> names;
  if (layer_names) {
    std::stringstream ss(layer_names);
    std::string n;
    while (std::getline(ss, n, '\n')) if (!n.empty()) names.push_back(n);
  } else {
    names = map.layers();
  }
  std::vector<ZipEntry> entries;
  auto add = [&](const std::string& name, const std::vector<char>& npy) {
    ZipEntry e{name, crc32(npy.data(), npy.size()), uint32_t(npy.size()), uint32_t(fs.tellp())};
    localHeader(fs, e);
    fs.write(npy.data(), npy.size());
    entries.push_back(e);
  };
  for (const auto& n : names) {
    if (!map.exists(n)) continue;  // "does not exist, skipping"
    add(n + ".npy", npyArray(map.get(n).data(), map.rows(), map.cols()));
  }
  add("meta.npy", npyString(metaJson(map, frame_id ? frame_id : "")));
  const uint32_t cd_offset = uint32_t(fs.tellp());
  for (const auto& e : entries) centralHeader(fs, e);
  const uint32_t cd_size = uint32_t(fs.tellp()) - cd_offset;
  w32(fs, 0x06054b50); w16(fs, 0); w16(fs, 0); w16(fs, uint16_t(entries.size()));
  w16(fs, uint16_t(entries.size())); w32(fs, cd_size); w32(fs, cd_offset); w16(fs, 0);
  return fs.fail() ? 0 : 1;
}

// loadNpz(filename, map): re-creates the geometry and layers of `map`; frame id copied into
// frame_out (cap bytes).  Returns 1 / 0.
int orc_load_npz(void* mp, const char* filename, char* frame_out, int cap) {
  auto& map = *static_cast<ElevationMap*>(mp);
  std::ifstream fs(filename, std::ios::binary);
  if (!fs.is_open()) return 0;
  struct Entry { std::string name; std::vector<char> data; };
  std::vector<Entry> entries;
  while (entries.size() < 1000) {
    const uint32_t sig = r32(fs);
    if (fs.fail() || sig != 0x04034b50) break;
    fs.ignore(14);
    r32(fs);
    const uint32_t usize = r32(fs);
    const uint16_t nlen = r16(fs), xlen = r16(fs);
    if (fs.fail()) break;
    if (nlen > 4096 || usize > 400000000u) return 0;
    std::string name(nlen, '\0');
    fs.read(&name[0], nlen);
    fs.ignore(xlen);
    std::vector<char> data(usize);
    fs.read(data.data(), usize);
    if (fs.fail()) return 0;
    entries.push_back({std::move(name), std::move(data)});
  }
  if (entries.empty()) return 0;
  std::string meta;
  bool found = false;
  for (const auto& e : entries) {
    if (e.name != "meta.npy") continue;
    NpyInfo info;
    if (!parseNpy(e.data.data(), e.data.size(), info) || !info.is_string) return 0;
    if (info.off + info.slen > e.data.size()) return 0;
    meta.assign(e.data.data() + info.off, info.slen);
    found = true;
    break;
  }
  if (!found) return 0;
  float version = 0, resolution = 0, px = 0, py = 0;
  int rows = 0, cols = 0, sx = 0, sy = 0;
  std::string frame;
  if (jsonFloat(meta, "version", version) && int(version) > 1) return 0;
  auto tof = [](const std::string& s) { return std::stof(s); };
  auto toi = [](const std::string& s) { return std::stoi(s); };
  if (!jsonFloat(meta, "resolution", resolution) || !jsonPair(meta, "position", px, py, tof) ||
      !jsonPair(meta, "size", rows, cols, toi))
    return 0;
  jsonString(meta, "frame_id", frame);
  jsonPair(meta, "start_index", sx, sy, toi);
  if (rows <= 0 || cols <= 0 || resolution <= 0) return 0;
  const float lx = resolution * rows, ly = resolution * cols;
  map.setGeometry(lx, ly, resolution);
  map.setPosition(px, py);
  map.setStartIndex(Index{sx, sy});
  if (frame_out && cap > 0) std::snprintf(frame_out, cap, "%s", frame.c_str());
  const size_t expected = size_t(rows) * cols * 4;
  int loaded = 0;
  for (const auto& e : entries) {
    if (e.name == "meta.npy") continue;
    if (e.name.size() <= 4 || e.name.substr(e.name.size() - 4) != ".npy") continue;
    const std::string lname = e.name.substr(0, e.name.size() - 4);
    NpyInfo info;
    if (!parseNpy(e.data.data(), e.data.size(), info) || !info.is_float) continue;
    if (info.rows != rows || info.cols != cols) continue;
    if (info.off + expected > e.data.size()) continue;
    if (!map.exists(lname)) map.add(lname);
    std::memcpy(map.get(lname).data(), e.data.data() + info.off, expected);
    ++loaded;
  }
  return loaded > 0 ? 1 : 0;
}

}  // extern "C"
