"""fastdem_b200 — B200-native implementation of FastDEM's per-scan integrate() path.

The product is libfastdem_b200.so (CUDA for sm_100a behind the C-ABI in
include/fastdem_b200.h).  This package is the host-side mirror of the reference's
interface for that path (api.py) plus the synthetic workloads (synthetic.py).  Importing
the package needs no GPU; using a map does, and fails loudly without one."""
from .capi import (EST_KALMAN, EST_P2QUANTILE, MODE_GLOBAL, MODE_LOCAL, MOVE_CLEAR_ALL_LAYERS,
                   MOVE_CLEAR_BASIC_LAYERS, SENSOR_CONSTANT, SENSOR_LIDAR, SENSOR_RGBD, FdemConfig,
                   FdemError, FdemGeometry, FdemScanStats, default_config, load_library)
from .api import (Config, ElevationMap, ElevationMapping, EstimationType, FastDEM, MappingMode,
                  PointCloud, PointCloud2, SensorType, applyFeatureExtraction, applyInpainting, applyRaycasting, applySpatialSmoothing,
                  applyUncertaintyFusion, layer, toPointCloud2, voxelGridAny)

from . import io_npz as io  # fastdem::io::{saveNpz, loadNpz}
from .config_yaml import loadConfig, parseConfig

__all__ = [
    "io", "loadConfig", "parseConfig",
    "Config", "ElevationMap", "ElevationMapping", "EstimationType", "FastDEM", "MappingMode",
    "PointCloud", "PointCloud2", "SensorType", "applyFeatureExtraction", "applyInpainting", "applyRaycasting",
    "applySpatialSmoothing", "applyUncertaintyFusion", "layer", "toPointCloud2", "voxelGridAny",
    "FdemConfig", "FdemError", "FdemGeometry", "FdemScanStats", "default_config", "load_library",
]
__version__ = "0.1.0"
