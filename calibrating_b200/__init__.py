"""calibrating_b200 -- B200-native engine for the `Stereo.get_depth` hot path of DIYer22/calibrating.

Public surface mirrors the reference's names for this path:
    Stereo, Cam                      (calibrating/stereo_camera.py, camera.py: load/dump schema only)
    MetaStereoMatching, SemiGlobalBlockMatching   (calibrating/stereo_matching.py:10-70)
    depth_to_point_cloud, point_cloud_to_depth    (calibrating/utils.py:213-288)
    interpolate_uvzs, interpolate_sparse2d, MatchingByBoard, FeatureMatchingAsStereoMatching   (utils.py:347-411, stereo_matching.py:73-142)
    StereoSGBM_create                (keyword-compatible with cv2.StereoSGBM_create, MODE_SGBM / MODE_HH / MODE_HH4)
Everything numeric runs in libb2s.so (hand-written sm_100a CUDA behind the C-ABI of include/b2s.h).
"""
from .stereo_matching import (COST_BT, COST_CENSUS, MODE_HH, MODE_HH4, MODE_SGBM, B200StereoMatching, FeatureMatchingAsStereoMatching, MatchingByBoard,
                              MetaStereoMatching, SemiGlobalBlockMatching, StereoSGBM, StereoSGBM_create)
from .stereo_camera import Cam, Stereo
from .utils import depth_to_point_cloud, interpolate_sparse2d, interpolate_uvzs, point_cloud_to_arr2d, point_cloud_to_depth

__all__ = ["Stereo", "Cam", "MetaStereoMatching", "SemiGlobalBlockMatching", "B200StereoMatching", "StereoSGBM",
           "StereoSGBM_create", "MatchingByBoard", "FeatureMatchingAsStereoMatching", "interpolate_uvzs", "interpolate_sparse2d", "depth_to_point_cloud", "point_cloud_to_depth", "point_cloud_to_arr2d", "MODE_SGBM", "MODE_HH", "MODE_HH4", "COST_BT", "COST_CENSUS"]
