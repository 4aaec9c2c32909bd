// elevation_map.hpp — fastdem::ElevationMap (fastdem/include/fastdem/elevation_map.hpp:65-177)
// backed by a device-resident map behind the C-ABI.  Same method names and meaning; layers live
// in HBM, so get(layer) returns a host COPY (nanogrid::Matrix, column-major) and at()/setAt()
// move single cells — the one visible difference from the reference's in-place Eigen matrices.
#pragma once

#include <cmath>
#include <initializer_list>
#include <stdexcept>
#include <string>
#include <vector>

#include "fastdem/compat.hpp"
#include "fastdem_b200.h"

namespace fastdem {

namespace layer {
constexpr auto elevation = "elevation";
constexpr auto elevation_min = "elevation_min";
constexpr auto elevation_max = "elevation_max";
constexpr auto variance = "variance";
constexpr auto n_points = "n_points";
constexpr auto upper_bound = "upper_bound";
constexpr auto lower_bound = "lower_bound";
constexpr auto obstacle = "obstacle";
constexpr auto intensity = "intensity";
constexpr auto color = "color";
constexpr auto kalman_p = "_kalman_p";
constexpr auto sample_mean = "_sample_mean";
constexpr auto sample_m2 = "_sample_m2";
constexpr auto ghost_removal = "ghost_removal";
constexpr auto raycasting = "raycasting";
constexpr auto visibility_logodds = "_visibility_logodds";
inline bool isInternal(const std::string& name) { return !name.empty() && name[0] == '_'; }
}  // namespace layer

// the reference logs + returns bool on the hot path and throws only from config / geometry
// misuse; a failing C-ABI call (no GPU, CUDA error) surfaces as this exception
struct DeviceError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
inline void check(fdem_status s) {
  if (s != FDEM_OK) throw DeviceError(std::string(fdem_status_string(s)) + ": " + fdem_last_error());
}

class ElevationMap {
 public:
  ElevationMap() = default;
  ElevationMap(float width, float height, float resolution, const std::string& frame_id,
               int device = 0) : device_(device) {
    setGeometry(width, height, resolution);
    setFrameId(frame_id);
  }
  ~ElevationMap() { if (h_) fdem_map_destroy(h_); }
  ElevationMap(const ElevationMap&) = delete;
  ElevationMap& operator=(const ElevationMap&) = delete;

  // nanoGrid setGeometry + clearAll (elevation_map.hpp:112-116).  On an existing map the resize
  // happens IN PLACE: the handle — and every FastDEM holding this map by reference — stays valid.
  void setGeometry(float width, float height, float resolution) {
    if (h_) check(fdem_map_set_geometry(h_, width, height, resolution));
    else check(fdem_map_create(width, height, resolution, device_, nullptr, &h_));
  }
  bool isInitialized() const { return h_ != nullptr; }
  fdem_map* handle() const { return h_; }

  // nanogrid::GridMap geometry
  nanogrid::Size getSize() const { auto g = geom(); return nanogrid::Size(g.rows, g.cols); }
  double getResolution() const { return geom().resolution; }
  nanogrid::Length getLength() const { auto g = geom(); return nanogrid::Length(g.length[0], g.length[1]); }
  nanogrid::Position getPosition() const { auto g = geom(); return nanogrid::Position(g.position[0], g.position[1]); }
  nanogrid::Index getStartIndex() const { auto g = geom(); return nanogrid::Index(g.start_index[0], g.start_index[1]); }
  void setPosition(const nanogrid::Position& p) { check(fdem_map_set_position(h_, p(0), p(1))); }
  void setStartIndex(const nanogrid::Index& i) { check(fdem_map_set_start_index(h_, i(0), i(1))); }
  const std::string& getFrameId() const { return frame_id_; }
  void setFrameId(const std::string& f) { frame_id_ = f; }

  bool isInside(const nanogrid::Position& p) const {
    int32_t in = 0;
    check(fdem_map_is_inside(h_, p(0), p(1), &in));
    return in != 0;
  }
  bool getIndex(const nanogrid::Position& p, nanogrid::Index& idx) const {
    int32_t r = 0, c = 0, in = 0;
    check(fdem_map_get_index(h_, p(0), p(1), &r, &c, &in));
    idx = nanogrid::Index(r, c);
    return in != 0;
  }
  bool getPosition(const nanogrid::Index& idx, nanogrid::Position& p) const {
    double x = 0, y = 0;
    check(fdem_map_get_cell_position(h_, idx(0), idx(1), &x, &y));
    p = nanogrid::Position(x, y);
    return true;
  }
  bool move(const nanogrid::Position& p) {
    int32_t moved = 0;
    check(fdem_map_move(h_, p(0), p(1), FDEM_MOVE_CLEAR_ALL_LAYERS, &moved));
    return moved != 0;
  }

  // layers
  bool exists(const std::string& name) const {
    int32_t e = 0;
    check(fdem_map_layer_exists(h_, name.c_str(), &e));
    return e != 0;
  }
  void add(const std::string& name, float fill = NAN) { check(fdem_map_layer_add(h_, name.c_str(), fill)); }
  std::vector<std::string> getLayers() const {
    int32_t n = 0;
    check(fdem_map_layer_count(h_, &n));
    std::vector<std::string> out;
    char buf[128];
    for (int i = 0; i < n; ++i) { check(fdem_map_layer_name(h_, i, buf, sizeof(buf))); out.emplace_back(buf); }
    return out;
  }
  nanogrid::Matrix get(const std::string& name) const {  // host copy
    auto g = geom();
    nanogrid::Matrix m(g.row_end - g.row_begin, g.cols);
    check(fdem_map_layer_download(h_, name.c_str(), m.data()));
    return m;
  }
  void set(const std::string& name, const nanogrid::Matrix& m) { check(fdem_map_layer_upload(h_, name.c_str(), m.data())); }
  float at(const std::string& name, const nanogrid::Index& idx) const {
    float v = NAN;
    check(fdem_map_cell_get(h_, name.c_str(), idx(0), idx(1), &v));
    return v;
  }
  // GridMap::at returns float&: `map.at(layer, idx) = v;` and `float v = map.at(layer, idx);` both
  // work through this proxy (the cell lives in HBM; a read or a write is one small transfer)
  class CellRef {
   public:
    operator float() const { return static_cast<const ElevationMap*>(map_)->at(name_, idx_); }  // the const overload: a read
    CellRef& operator=(float v) { map_->setAt(name_, idx_, v); return *this; }
    CellRef& operator=(const CellRef& o) { return *this = static_cast<float>(o); }
    CellRef& operator+=(float v) { return *this = static_cast<float>(*this) + v; }
    CellRef& operator-=(float v) { return *this = static_cast<float>(*this) - v; }

   private:
    friend class ElevationMap;
    CellRef(ElevationMap* m, std::string n, nanogrid::Index i) : map_(m), name_(std::move(n)), idx_(i) {}
    ElevationMap* map_;
    std::string name_;
    nanogrid::Index idx_;
  };
  CellRef at(const std::string& name, const nanogrid::Index& idx) { return CellRef(this, name, idx); }
  void setAt(const std::string& name, const nanogrid::Index& idx, float v) {
    check(fdem_map_cell_set(h_, name.c_str(), idx(0), idx(1), v));
  }
  void clear(const std::string& name) { check(fdem_map_clear(h_, name.c_str())); }
  void clearAll() { check(fdem_map_clear_all(h_)); }

  // ElevationMap conveniences (elevation_map.hpp:118-177)
  bool isEmpty() const { int32_t e = 0; check(fdem_map_is_empty(h_, &e)); return e != 0; }
  bool isEmptyAt(const nanogrid::Index& idx) const { return std::isnan(at(layer::elevation, idx)); }
  void clearAt(const nanogrid::Index& idx) { check(fdem_map_clear_at(h_, idx(0), idx(1))); }
  float elevationAt(const nanogrid::Position& p) const {
    nanogrid::Index idx;
    if (!getIndex(p, idx)) return NAN;
    return at(layer::elevation, idx);
  }
  float elevationAt(const nanogrid::Index& idx) const { return at(layer::elevation, idx); }
  bool hasElevationAt(const nanogrid::Position& p) const { return std::isfinite(elevationAt(p)); }
  bool hasElevationAt(const nanogrid::Index& idx) const { return std::isfinite(elevationAt(idx)); }
  // ElevationMap::snapshot (elevation_map.hpp:161-177): a second map with the same geometry,
  // position, start index and frame holding copies of the named layers (missing names skipped).
  // Device-to-host-to-device here: the post-process timer of the ROS node calls it at 1-2 Hz.
  void snapshot(ElevationMap& snap, std::initializer_list<std::string> layers) const {
    auto g = geom();
    snap.device_ = device_;
    snap.setGeometry(static_cast<float>(g.length[0]), static_cast<float>(g.length[1]), static_cast<float>(g.resolution));
    snap.setFrameId(getFrameId());
    snap.setPosition(getPosition());
    snap.setStartIndex(getStartIndex());
    for (const auto& name : layers) {
      if (!exists(name)) continue;
      if (!snap.exists(name)) snap.add(name);
      snap.set(name, get(name));
    }
  }
  // move construction only (the handle is unique), so `auto snap = map.snapshot({...})` works
  ElevationMap(ElevationMap&& o) noexcept : h_(o.h_), device_(o.device_), frame_id_(std::move(o.frame_id_)) { o.h_ = nullptr; }
  ElevationMap snapshot(std::initializer_list<std::string> layers) const {
    ElevationMap snap;
    snapshot(snap, layers);
    return snap;
  }

 private:
  fdem_geometry geom() const {
    fdem_geometry g{};
    if (!h_) return g;
    check(fdem_map_get_geometry(h_, &g));
    return g;
  }
  fdem_map* h_ = nullptr;
  int device_ = 0;
  std::string frame_id_;
};

}  // namespace fastdem
