// config_io.hpp — fastdem::loadConfig / parseConfig for C++ callers
// (fastdem/src/config_fastdem.cpp:57-277, fastdem/config/default.yaml).
//
// The reference reads its YAML through yaml-cpp.  Its config files use a small subset of YAML —
// nested block mappings of scalars, comments, optional quotes — so this header carries a parser
// for exactly that subset (no sequences, anchors or flow collections) instead of a dependency.
// Semantics follow the reference: keys that are absent keep their defaults; unknown enum strings
// fall back with a warning (detail::parse*), detail::validate throws std::invalid_argument on the
// two fatal inconsistencies and clamps the rest with a warning (fdem_config_validate implements
// the rules once for every host language); a file that cannot be read is std::runtime_error.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "fastdem/config.hpp"

namespace fastdem {
namespace detail {

// flat view of a block-mapping document: "mapping.kalman.min_variance" -> "0.0001"
inline std::map<std::string, std::string> parseYamlScalars(std::istream& in) {
  std::map<std::string, std::string> out;
  std::vector<std::pair<int, std::string>> stack;  // (indent, key) of the open mappings
  std::string line;
  while (std::getline(in, line)) {
    // strip comments (a '#' outside quotes, at line start or after whitespace)
    bool sq = false, dq = false;
    for (size_t i = 0; i < line.size(); ++i) {
      const char ch = line[i];
      if (ch == '\'' && !dq) sq = !sq;
      else if (ch == '"' && !sq) dq = !dq;
      else if (ch == '#' && !sq && !dq && (i == 0 || line[i - 1] == ' ' || line[i - 1] == '\t')) {
        line.erase(i);
        break;
      }
    }
    size_t end = line.find_last_not_of(" \t\r\n");
    if (end == std::string::npos) continue;
    line.erase(end + 1);
    const size_t indent = line.find_first_not_of(' ');
    if (indent == std::string::npos || line.compare(indent, 3, "---") == 0) continue;
    const size_t colon = line.find(':', indent);
    if (colon == std::string::npos) continue;  // not a mapping entry: outside the subset, ignored
    std::string key = line.substr(indent, colon - indent);
    while (!key.empty() && (key.back() == ' ' || key.back() == '\t')) key.pop_back();
    if (key.size() >= 2 && (key.front() == '"' || key.front() == '\'')) key = key.substr(1, key.size() - 2);
    std::string val = colon + 1 < line.size() ? line.substr(colon + 1) : "";
    const size_t vb = val.find_first_not_of(" \t");
    val = vb == std::string::npos ? "" : val.substr(vb);
    if (val.size() >= 2 && ((val.front() == '"' && val.back() == '"') || (val.front() == '\'' && val.back() == '\'')))
      val = val.substr(1, val.size() - 2);
    while (!stack.empty() && stack.back().first >= static_cast<int>(indent)) stack.pop_back();
    std::string path;
    for (const auto& s : stack) path += s.second + ".";
    path += key;
    if (val.empty()) stack.emplace_back(static_cast<int>(indent), key);  // opens a nested mapping
    else out[path] = val;
  }
  return out;
}

inline void warn(const std::string& msg) { std::fprintf(stderr, "[Config] %s\n", msg.c_str()); }

struct Scalars {
  std::map<std::string, std::string> kv;
  bool has(const std::string& k) const { return kv.count(k) != 0; }
  void load(const std::string& k, float& v) const {
    auto it = kv.find(k);
    if (it == kv.end()) return;
    char* e = nullptr;
    const double d = std::strtod(it->second.c_str(), &e);
    if (e == it->second.c_str()) throw std::runtime_error("bad float for '" + k + "': " + it->second);
    v = static_cast<float>(d);
  }
  void load(const std::string& k, int& v) const {
    auto it = kv.find(k);
    if (it == kv.end()) return;
    char* e = nullptr;
    const long d = std::strtol(it->second.c_str(), &e, 10);
    if (e == it->second.c_str()) throw std::runtime_error("bad integer for '" + k + "': " + it->second);
    v = static_cast<int>(d);
  }
  void load(const std::string& k, bool& v) const {
    auto it = kv.find(k);
    if (it == kv.end()) return;
    const std::string& s = it->second;  // yaml-cpp's bool conversion: true/false, yes/no, on/off (any case)
    std::string l;
    for (char ch : s) l += static_cast<char>(std::tolower(static_cast<unsigned char>(ch)));
    if (l == "true" || l == "yes" || l == "on" || l == "y") v = true;
    else if (l == "false" || l == "no" || l == "off" || l == "n") v = false;
    else throw std::runtime_error("bad bool for '" + k + "': " + s);
  }
  void load(const std::string& k, std::string& v) const {
    auto it = kv.find(k);
    if (it != kv.end()) v = it->second;
  }
};

// detail::parse (config_fastdem.cpp:57-126)
inline Config parse(const Scalars& y) {
  Config cfg;
  auto& m = cfg.mapping;
  std::string mode, type, sensor;
  y.load("mapping.mode", mode);
  if (!mode.empty()) {
    if (mode == "local") m.mode = MappingMode::LOCAL;
    else if (mode == "global") m.mode = MappingMode::GLOBAL;
    else { warn("Unknown mapping mode '" + mode + "', defaulting to local"); m.mode = MappingMode::LOCAL; }
  }
  y.load("mapping.type", type);
  if (!type.empty()) {
    if (type == "kalman_filter") m.estimation_type = EstimationType::Kalman;
    else if (type == "p2_quantile") m.estimation_type = EstimationType::P2Quantile;
    else { warn("Unknown estimation type '" + type + "', defaulting to kalman_filter"); m.estimation_type = EstimationType::Kalman; }
  }
  y.load("mapping.kalman.min_variance", m.kalman.min_variance);
  y.load("mapping.kalman.max_variance", m.kalman.max_variance);
  y.load("mapping.kalman.process_noise", m.kalman.process_noise);
  y.load("mapping.p2.dn0", m.p2.dn0);
  y.load("mapping.p2.dn1", m.p2.dn1);
  y.load("mapping.p2.dn2", m.p2.dn2);
  y.load("mapping.p2.dn3", m.p2.dn3);
  y.load("mapping.p2.dn4", m.p2.dn4);
  y.load("mapping.p2.elevation_marker", m.p2.elevation_marker);
  y.load("mapping.p2.max_sample_count", m.p2.max_sample_count);
  y.load("point_filter.z_min", cfg.point_filter.z_min);
  y.load("point_filter.z_max", cfg.point_filter.z_max);
  y.load("point_filter.range_min", cfg.point_filter.range_min);
  y.load("point_filter.range_max", cfg.point_filter.range_max);
  y.load("raycasting.enabled", cfg.raycasting.enabled);
  y.load("raycasting.height_conflict_threshold", cfg.raycasting.height_conflict_threshold);
  y.load("raycasting.log_odds_observed", cfg.raycasting.log_odds_observed);
  y.load("raycasting.log_odds_ghost", cfg.raycasting.log_odds_ghost);
  y.load("raycasting.log_odds_max", cfg.raycasting.log_odds_max);
  y.load("raycasting.clear_threshold", cfg.raycasting.clear_threshold);
  y.load("sensor_model.type", sensor);
  if (!sensor.empty()) {
    if (sensor == "lidar" || sensor == "laser") cfg.sensor_model.type = SensorType::LiDAR;
    else if (sensor == "rgbd") cfg.sensor_model.type = SensorType::RGBD;
    else if (sensor == "constant" || sensor == "none") cfg.sensor_model.type = SensorType::Constant;
    else { warn("Unknown sensor_model.type '" + sensor + "', defaulting to LiDAR"); cfg.sensor_model.type = SensorType::LiDAR; }
  }
  y.load("sensor_model.lidar.range_noise", cfg.sensor_model.lidar.range_noise);
  y.load("sensor_model.lidar.angular_noise", cfg.sensor_model.lidar.angular_noise);
  y.load("sensor_model.rgbd.normal_a", cfg.sensor_model.rgbd.normal_a);
  y.load("sensor_model.rgbd.normal_b", cfg.sensor_model.rgbd.normal_b);
  y.load("sensor_model.rgbd.normal_c", cfg.sensor_model.rgbd.normal_c);
  y.load("sensor_model.rgbd.lateral_factor", cfg.sensor_model.rgbd.lateral_factor);
  y.load("sensor_model.constant.uncertainty", cfg.sensor_model.constant.uncertainty);
  return cfg;
}

// the inverse of toAbi, for the fields validate() may clamp
inline void fromAbi(const fdem_config& a, Config& c) {
  c.mapping.kalman.min_variance = a.kalman_min_variance;
  c.mapping.kalman.max_variance = a.kalman_max_variance;
  c.mapping.kalman.process_noise = a.kalman_process_noise;
  c.mapping.p2.dn0 = a.p2_dn[0]; c.mapping.p2.dn1 = a.p2_dn[1]; c.mapping.p2.dn2 = a.p2_dn[2];
  c.mapping.p2.dn3 = a.p2_dn[3]; c.mapping.p2.dn4 = a.p2_dn[4];
  c.mapping.p2.elevation_marker = a.p2_elevation_marker;
  c.raycasting.height_conflict_threshold = a.rc_height_conflict_threshold;
  c.raycasting.log_odds_observed = a.rc_log_odds_observed;
  c.raycasting.log_odds_ghost = a.rc_log_odds_ghost;
  c.raycasting.log_odds_max = a.rc_log_odds_max;
  c.raycasting.clear_threshold = a.rc_clear_threshold;
  c.sensor_model.lidar.range_noise = a.lidar_range_noise;
  c.sensor_model.lidar.angular_noise = a.lidar_angular_noise;
  c.sensor_model.constant.uncertainty = a.constant_uncertainty;
  c.sensor_model.rgbd.normal_a = a.rgbd_normal_a; c.sensor_model.rgbd.normal_b = a.rgbd_normal_b;
  c.sensor_model.rgbd.normal_c = a.rgbd_normal_c; c.sensor_model.rgbd.lateral_factor = a.rgbd_lateral_factor;
}

// detail::validate (config_fastdem.cpp:128-260)
inline void validate(Config& cfg) {
  fdem_config a = toAbi(cfg);
  int32_t clamped = 0;
  if (fdem_config_validate(&a, &clamped) != FDEM_OK) throw std::invalid_argument(fdem_last_error());
  if (clamped) warn(std::to_string(clamped) + " value(s) out of range, clamped (see config_fastdem.cpp:128-260)");
  fromAbi(a, cfg);
}

}  // namespace detail

// fastdem::parseConfig(YAML::Node) for a YAML document held in a string
inline Config parseConfig(const std::string& yaml_text) {
  std::istringstream in(yaml_text);
  detail::Scalars y{detail::parseYamlScalars(in)};
  Config cfg = detail::parse(y);
  detail::validate(cfg);
  return cfg;
}

// fastdem::loadConfig(path) (config_fastdem.cpp:270-277)
inline Config loadConfig(const std::string& path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("Failed to load config: " + path);
  try {
    detail::Scalars y{detail::parseYamlScalars(in)};
    Config cfg = detail::parse(y);
    detail::validate(cfg);
    return cfg;
  } catch (const std::invalid_argument&) {
    throw;
  } catch (const std::exception& e) {
    throw std::runtime_error("Failed to load config: " + path + " - " + e.what());
  }
}

}  // namespace fastdem
