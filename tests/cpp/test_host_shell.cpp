// test_host_shell.cpp — a few of the reference's gtest cases (fastdem/tests/
// test_fastdem_integration.cpp, test_elevation_map.cpp, test_postprocess.cpp), written against the
// C++ host shell exactly as they are written against the reference.  Built and run by
// tests/test_gpu_cpp_shell.py on the GPU box.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "fastdem/fastdem.hpp"

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
  if (g_failures == 0) std::printf("ALL C++ SHELL TESTS PASSED\n");
  return g_failures == 0 ? 0 : 1;
}
