// test_host_shell.cpp — a few of the reference's gtest cases (fastdem/tests/
// test_fastdem_integration.cpp, test_elevation_map.cpp, test_postprocess.cpp), written against the
// C++ host shell exactly as they are written against the reference.  Built and run by
// tests/test_gpu_cpp_shell.py on the GPU box.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "fastdem/config_io.hpp"
#include "fastdem/fastdem.hpp"
#include "fastdem/io_npz.hpp"

using namespace fastdem;

static int g_failures = 0;
#define EXPECT_TRUE(c) do { if (!(c)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); ++g_failures; } } while (0)
#define EXPECT_FALSE(c) EXPECT_TRUE(!(c))
#define EXPECT_NEAR(a, b, tol) EXPECT_TRUE(std::fabs((a) - (b)) <= (tol))

static PointCloud makeGroundCloud(float height, int grid_half = 3, float spacing = 0.3f) {
  PointCloud cloud;  // test_fastdem_integration.cpp:32-41
  for (int i = -grid_half; i <= grid_half; ++i)
    for (int j = -grid_half; j <= grid_half; ++j) cloud.add(i * spacing, j * spacing, height);
  return cloud;
}

int main() {
  const Eigen::Isometry3d I = Eigen::Isometry3d::Identity();
  {  // IntegrateUpdatesElevation (:45-59)
    ElevationMap map;
    map.setGeometry(10.0f, 10.0f, 0.5f);
    EXPECT_TRUE(map.isInitialized());
    EXPECT_TRUE(map.isEmpty());
    FastDEM mapper(map);
    mapper.setHeightFilter(-2.0f, 5.0f).setRangeFilter(0.0f, 20.0f).setSensorModel(SensorType::Constant)
        .setEstimatorType(EstimationType::Kalman);
    EXPECT_TRUE(mapper.integrate(makeGroundCloud(1.0f), I, I));
    nanogrid::Position center(0.0, 0.0);
    EXPECT_TRUE(map.hasElevationAt(center));
    EXPECT_NEAR(map.elevationAt(center), 1.0f, 0.1f);
    EXPECT_TRUE(map.exists(layer::kalman_p));
  }
  {  // EmptyCloudIsNoOp / IntegrateReturnsFalseOnEmpty / ...WhenAllFiltered (:61-69, :365-378)
    ElevationMap map(10.0f, 10.0f, 0.5f, "map");
    FastDEM mapper(map);
    PointCloud empty;
    EXPECT_FALSE(mapper.integrate(empty, I, I));
    EXPECT_TRUE(map.isEmpty());
    mapper.setHeightFilter(100.0f, 200.0f);
    EXPECT_FALSE(mapper.integrate(makeGroundCloud(1.0f), I, I));
    EXPECT_TRUE(map.isEmpty());
  }
  {  // LocalModeFollowsRobot (:198-215) + P2 after 6 scans (:159-175)
    ElevationMap map(10.0f, 10.0f, 0.5f, "map");
    FastDEM mapper(map);
    mapper.setMappingMode(MappingMode::LOCAL).setHeightFilter(-5.0f, 15.0f).setSensorModel(SensorType::Constant)
        .setEstimatorType(EstimationType::P2Quantile);
    for (int i = 0; i < 6; ++i) mapper.integrate(makeGroundCloud(1.0f + i * 0.01f), I, I);
    EXPECT_NEAR(map.elevationAt(nanogrid::Position(0.0, 0.0)), 1.0f, 0.2f);
    Eigen::Isometry3d T = Eigen::Isometry3d::Identity();
    T.translation().x() = 100.0;
    mapper.integrate(makeGroundCloud(2.0f), I, T);
    EXPECT_FALSE(map.isInside(nanogrid::Position(0.0, 0.0)));
  }
  {  // callbacks (:320-353)
    ElevationMap map(10.0f, 10.0f, 0.5f, "map");
    FastDEM mapper(map);
    mapper.setHeightFilter(-5.0f, 15.0f).setSensorModel(SensorType::Constant);
    size_t pre = 0, ras = 0;
    mapper.onScanPreprocessed([&](const PointCloud& c) { pre = c.size(); });
    mapper.onScanRasterized([&](const PointCloud& c) { ras = c.size(); });
    mapper.integrate(makeGroundCloud(1.0f), I, I);
    EXPECT_TRUE(pre == 49);
    EXPECT_TRUE(ras > 0 && ras <= 49);
  }
  {  // RaycastingClearsGhostCell (test_postprocess.cpp:92-115)
    ElevationMap map(10.0f, 10.0f, 0.5f, "map");
    nanogrid::Index ghost_idx;
    EXPECT_TRUE(map.getIndex(nanogrid::Position(2.0, 0.0), ghost_idx));
    map.setAt(layer::elevation, ghost_idx, 10.0f);
    PointCloud cloud;
    cloud.add(4.0f, 0.0f, 0.0f);
    config::Raycasting cfg;
    cfg.enabled = true;
    cfg.height_conflict_threshold = 0.05f;
    cfg.log_odds_ghost = 0.5f;
    cfg.clear_threshold = -0.4f;
    applyRaycasting(map, cloud, Eigen::Vector3f(0.0f, 0.0f, 5.0f), cfg);
    EXPECT_TRUE(std::isnan(map.at(layer::elevation, ghost_idx)));
    EXPECT_NEAR(map.at(layer::ghost_removal, ghost_idx), 1.0f, 0.0f);
  }
  {  // UncertaintyFusionComputesBounds + FeatureExtractionFlatPlane (test_postprocess.cpp:193-225, 285-297)
    ElevationMap map(10.0f, 10.0f, 0.5f, "map");
    map.add(layer::upper_bound, NAN);
    map.add(layer::lower_bound, NAN);
    const nanogrid::Index center(10, 10);
    for (int dr = -1; dr <= 1; ++dr)
      for (int dc = -1; dc <= 1; ++dc) {
        const nanogrid::Index idx(center(0) + dr, center(1) + dc);
        const float h = 1.0f + 0.1f * dr;
        map.setAt(layer::elevation, idx, h);
        map.setAt(layer::upper_bound, idx, h + 0.2f);
        map.setAt(layer::lower_bound, idx, h - 0.2f);
      }
    config::UncertaintyFusion uf;
    uf.enabled = true;
    uf.search_radius = 0.6f;
    uf.spatial_sigma = 0.3f;
    uf.min_valid_neighbors = 1;
    applyUncertaintyFusion(map, uf);
    const float upper = map.at(layer::upper_bound, center), lower = map.at(layer::lower_bound, center);
    EXPECT_TRUE(std::isfinite(upper) && std::isfinite(lower) && upper > lower);

    ElevationMap flat(10.0f, 10.0f, 0.5f, "map");
    nanogrid::Matrix ones(20, 20);
    for (int i = 0; i < 400; ++i) ones.data()[i] = 1.0f;
    flat.set(layer::elevation, ones);
    applyFeatureExtraction(flat, 0.6f, 4);
    EXPECT_TRUE(flat.exists("slope") && flat.exists("_normal_z"));
    EXPECT_NEAR(flat.at("slope", center), 0.0f, 1.0f);
    EXPECT_NEAR(flat.at("roughness", center), 0.0f, 0.001f);
    EXPECT_NEAR(flat.at("_normal_z", center), 1.0f, 0.01f);
    applySpatialSmoothing(flat, layer::elevation, 3, 5);
    EXPECT_NEAR(flat.at(layer::elevation, center), 1.0f, 0.0f);
  }
  {  // ElevationMap index round trip + clearAt (test_elevation_map.cpp:40-61)
    ElevationMap map(10.0f, 10.0f, 0.5f, "world");
    nanogrid::Position pos(1.0, 1.0);
    nanogrid::Index idx;
    EXPECT_TRUE(map.getIndex(pos, idx));
    map.setAt(layer::elevation, idx, 1.5f);
    EXPECT_TRUE(map.hasElevationAt(pos));
    map.clearAt(idx);
    EXPECT_FALSE(map.hasElevationAt(pos));
    EXPECT_TRUE(std::isnan(map.elevationAt(nanogrid::Position(100.0, 100.0))));
    EXPECT_TRUE(map.getFrameId() == "world");
  }
  {  // setSensorModel(std::unique_ptr<SensorModel>) (fastdem.hpp:80): a model passed as an object
     // gives the same map as the same model selected by enum; a user subclass is honoured
    struct Doubled : LiDARSensorModel {
      Eigen::Matrix3f computeCovariance(const Eigen::Vector3f& p) const override {
        return LiDARSensorModel::computeCovariance(p) * 2.0f;
      }
    };
    ElevationMap a(10.0f, 10.0f, 0.5f, "map"), b(10.0f, 10.0f, 0.5f, "map"), c(10.0f, 10.0f, 0.5f, "map");
    FastDEM ma(a), mb(b), mc(c);
    for (FastDEM* m : {&ma, &mb, &mc}) m->setHeightFilter(-5.0f, 15.0f).setMappingMode(MappingMode::GLOBAL);
    ma.setSensorModel(SensorType::LiDAR);
    mb.setSensorModel(std::make_unique<LiDARSensorModel>());
    mc.setSensorModel(std::make_unique<Doubled>());
    Eigen::Isometry3d Ts = Eigen::Isometry3d::Identity();
    Ts.translation().z() = 1.5;
    for (int k = 0; k < 3; ++k) {
      EXPECT_TRUE(ma.integrate(makeGroundCloud(-1.0f + 0.01f * k), Ts, I));
      EXPECT_TRUE(mb.integrate(makeGroundCloud(-1.0f + 0.01f * k), Ts, I));
      EXPECT_TRUE(mc.integrate(makeGroundCloud(-1.0f + 0.01f * k), Ts, I));
    }
    const nanogrid::Matrix pa = a.get(layer::kalman_p), pb = b.get(layer::kalman_p), pc = c.get(layer::kalman_p);
    const nanogrid::Matrix ea = a.get(layer::elevation), eb = b.get(layer::elevation);
    int finite = 0, larger = 0;
    for (size_t i = 0; i < pa.size(); ++i) {
      if (!std::isfinite(ea.data()[i])) { EXPECT_TRUE(!std::isfinite(eb.data()[i])); continue; }  // unobserved cell
      ++finite;
      EXPECT_TRUE(pa.data()[i] == pb.data()[i] && ea.data()[i] == eb.data()[i]);   // bit for bit
      if (pc.data()[i] > pa.data()[i]) ++larger;
    }
    EXPECT_TRUE(finite > 10 && larger > 0);
  }
  {  // onScanPreprocessed carries the covariance channel (fastdem.cpp:139-141, 181-187)
    ElevationMap map(10.0f, 10.0f, 0.5f, "map");
    FastDEM mapper(map);
    mapper.setHeightFilter(-5.0f, 15.0f).setSensorModel(SensorType::Constant);
    bool has_cov = false;
    float c22 = 0.0f;
    mapper.onScanPreprocessed([&](const PointCloud& c) { has_cov = c.hasCovariance(); c22 = c.covariance(0)(2, 2); });
    mapper.integrate(makeGroundCloud(1.0f), I, I);
    EXPECT_TRUE(has_cov);
    EXPECT_NEAR(c22, 0.03f * 0.03f, 1e-8f);   // config default sigma = 0.03 (config/sensor_model.hpp:35)
  }
  {  // ElevationMapping::update, the lower seam (test_dual_layer.cpp:66-83, 121-143)
    ElevationMap map(10.0f, 10.0f, 0.5f, "map");
    config::Mapping mc;
    mc.mode = MappingMode::GLOBAL;
    mc.kalman.min_variance = 0.0001f;
    mc.kalman.max_variance = 1.0f;
    auto mapping = createElevationMapping(map, mc);
    PointCloud cloud;
    cloud.add(0.1f, 0.1f, 0.0f);
    cloud.add(0.1f, 0.1f, 3.0f);
    auto obs = mapping->update(cloud, Eigen::Vector2d(0.0, 0.0));
    EXPECT_TRUE(obs.size() == 1);
    EXPECT_NEAR(obs[0].obs.min_z, 0.0f, 0.0f);
    nanogrid::Index idx;
    EXPECT_TRUE(map.getIndex(nanogrid::Position(0.1, 0.1), idx));
    EXPECT_NEAR(map.at(layer::elevation, idx), 0.0f, 0.1f);
    EXPECT_NEAR(map.at(layer::obstacle, idx), 3.0f, 0.0f);
    PointCloud second;
    second.add(0.1f, 0.1f, 0.05f);
    second.add(0.1f, 0.1f, 3.1f);
    mapping->update(second, Eigen::Vector2d(0.0, 0.0));
    EXPECT_TRUE(static_cast<float>(map.at(layer::obstacle, idx)) == 3.1f);
  }
  {  // at() as an lvalue + snapshot (elevation_map.hpp:161-177)
    ElevationMap map(10.0f, 10.0f, 0.5f, "odom");
    const nanogrid::Index idx(3, 4);
    map.at(layer::elevation, idx) = 2.5f;
    map.at(layer::elevation, idx) += 0.25f;
    const float v = map.at(layer::elevation, idx);
    EXPECT_TRUE(v == 2.75f);
    map.add("upper_bound", 1.0f);
    map.setPosition(nanogrid::Position(1.0, -2.0));
    ElevationMap snap = map.snapshot({layer::elevation, "upper_bound", "no_such_layer"});
    EXPECT_TRUE(snap.getFrameId() == "odom");
    EXPECT_TRUE(snap.getPosition()(0) == 1.0 && snap.getPosition()(1) == -2.0);
    EXPECT_TRUE(snap.exists("upper_bound") && !snap.exists("no_such_layer") && !snap.exists("variance"));
    EXPECT_TRUE(static_cast<float>(snap.at(layer::elevation, idx)) == 2.75f);
    map.at(layer::elevation, idx) = 9.0f;                       // the snapshot is a copy
    EXPECT_TRUE(static_cast<float>(snap.at(layer::elevation, idx)) == 2.75f);
  }
  {  // setGeometry on a live map keeps the FastDEM bound to it valid (what io::loadNpz relies on)
    ElevationMap map(10.0f, 10.0f, 0.5f, "map");
    FastDEM mapper(map);
    mapper.setHeightFilter(-5.0f, 15.0f).setSensorModel(SensorType::Constant);
    EXPECT_TRUE(mapper.integrate(makeGroundCloud(1.0f), I, I));
    map.setGeometry(6.0f, 8.0f, 0.25f);
    EXPECT_TRUE(map.getSize()(0) == 24 && map.getSize()(1) == 32 && map.isEmpty());
    EXPECT_TRUE(mapper.integrate(makeGroundCloud(1.0f), I, I));
    EXPECT_NEAR(map.elevationAt(nanogrid::Position(0.0, 0.0)), 1.0f, 0.1f);
  }
  {  // loadConfig / parseConfig (test_config.cpp: defaults, overrides, clamping, fatal errors)
    const Config d = parseConfig("");
    EXPECT_TRUE(d.mapping.mode == MappingMode::LOCAL && d.sensor_model.type == SensorType::LiDAR);
    const char* text =
        "# FastDEM Configuration\n"
        "mapping:\n"
        "  mode: \"global\"            # local | global\n"
        "  type: 'p2_quantile'\n"
        "  kalman:\n"
        "    min_variance: 0.0002\n"
        "    process_noise: -1.0      # clamped to 0\n"
        "  p2:\n"
        "    elevation_marker: 9      # clamped to 4\n"
        "point_filter:\n"
        "  z_min: -1.0\n"
        "  range_max: 20.0\n"
        "sensor_model:\n"
        "  type: rgbd\n"
        "  rgbd:\n"
        "    normal_c: 0.5\n"
        "raycasting:\n"
        "  enabled: true\n"
        "  clear_threshold: 0.5       # must be < 0: clamped to -1\n";
    const Config c = parseConfig(text);
    EXPECT_TRUE(c.mapping.mode == MappingMode::GLOBAL && c.mapping.estimation_type == EstimationType::P2Quantile);
    EXPECT_NEAR(c.mapping.kalman.min_variance, 0.0002f, 1e-9f);
    EXPECT_NEAR(c.mapping.kalman.process_noise, 0.0f, 0.0f);
    EXPECT_TRUE(c.mapping.p2.elevation_marker == 4);
    EXPECT_NEAR(c.point_filter.z_min, -1.0f, 0.0f);
    EXPECT_NEAR(c.point_filter.range_max, 20.0f, 0.0f);
    EXPECT_TRUE(c.sensor_model.type == SensorType::RGBD);
    EXPECT_NEAR(c.sensor_model.rgbd.normal_c, 0.5f, 0.0f);
    EXPECT_TRUE(c.raycasting.enabled);
    EXPECT_NEAR(c.raycasting.clear_threshold, -1.0f, 0.0f);
    bool threw = false;
    try { parseConfig("mapping:\n  kalman:\n    min_variance: 0.5\n    max_variance: 0.1\n"); }
    catch (const std::invalid_argument&) { threw = true; }
    EXPECT_TRUE(threw);
    threw = false;
    try { loadConfig("/nonexistent/fastdem.yaml"); } catch (const std::runtime_error&) { threw = true; }
    EXPECT_TRUE(threw);
  }
  {  // io::saveNpz / loadNpz round trip incl. start index (test_map_io.cpp:39-82), into a LIVE map
    ElevationMap map(10.0f, 10.0f, 0.5f, "odom");
    FastDEM mapper(map);
    mapper.setHeightFilter(-5.0f, 15.0f).setSensorModel(SensorType::Constant);
    Eigen::Isometry3d T = Eigen::Isometry3d::Identity();
    T.translation().x() = 1.3;           // LOCAL mode: the window moves, start index != 0
    mapper.integrate(makeGroundCloud(1.0f), I, T);
    const char* path = "/tmp/fdem_cpp_shell_test.npz";
    EXPECT_TRUE(io::saveNpz(path, map));
    ElevationMap other(4.0f, 4.0f, 1.0f, "x");
    FastDEM other_mapper(other);
    other_mapper.setHeightFilter(-5.0f, 15.0f).setSensorModel(SensorType::Constant);
    EXPECT_TRUE(io::loadNpz(path, other));
    EXPECT_TRUE(other.getSize()(0) == 20 && other.getSize()(1) == 20 && other.getFrameId() == "odom");
    EXPECT_TRUE(other.getStartIndex()(0) == map.getStartIndex()(0) && other.getStartIndex()(0) != 0);
    EXPECT_TRUE(other.getPosition()(0) == map.getPosition()(0));
    const nanogrid::Matrix a = map.get(layer::elevation), b = other.get(layer::elevation);
    const nanogrid::Matrix pa = map.get(layer::kalman_p), pb = other.get(layer::kalman_p);
    int same = 0;
    for (size_t i = 0; i < a.size(); ++i)
      if ((std::isnan(a.data()[i]) && std::isnan(b.data()[i])) || (a.data()[i] == b.data()[i] && pa.data()[i] == pb.data()[i])) ++same;
    EXPECT_TRUE(same == 400);
    EXPECT_TRUE(other_mapper.integrate(makeGroundCloud(1.0f), I, T));   // the bound FastDEM survived loadNpz
    EXPECT_FALSE(io::loadNpz("/nonexistent/map.npz", other));
  }
  if (g_failures == 0) std::printf("ALL C++ SHELL TESTS PASSED\n");
  return g_failures == 0 ? 0 : 1;
}
