"""Sensor tables the rasterizer's callers need (fixtures; values are the reference's:
src/torchbox3d/prototype/loader.py:62-129 == datasets/argoverse/constants.py:560-627 and :231-266)."""
import numpy as np

# laser number -> row (before the H - row - 1 flip) for the 64-beam AV2 range image
ROW_MAPPING_64 = np.array(
    [56, 22, 42, 28, 61, 30, 49, 36, 40, 32, 38, 45, 34, 26, 53, 59, 8, 1, 16, 20, 12, 5, 11, 15, 17, 9, 24, 6,
     13, 3, 19, 0, 7, 41, 21, 35, 2, 33, 14, 27, 23, 31, 25, 18, 29, 37, 10, 4, 55, 62, 47, 43, 51, 58, 52, 48,
     46, 54, 39, 57, 50, 60, 44, 63])

# laser number -> row for the 32-beam (single lidar) range image (datasets/argoverse/constants.py:453-488)
ROW_MAPPING_32 = np.array(
    [29, 15, 25, 18, 31, 19, 27, 22, 24, 20, 23, 26, 21, 17, 28, 30, 5, 1, 11, 14, 8, 3, 7, 10, 12, 6, 16, 4, 9, 2,
     13, 0])

# per-log laser renumbering used by correct_laser_numbers (converters/av2/utils.py:211-226)
LASER_MAPPING = np.array(
    [4, 15, 0, 14, 6, 11, 2, 8, 10, 7, 12, 9, 5, 3, 13, 26, 1, 19, 30, 24, 18, 23, 28, 20, 22, 25, 16, 27, 21,
     29, 17, 31])

# egovehicle -> up-lidar translation the reference hard-codes (datasets/argoverse/av2.py:162)
LIDAR_OFFSET = np.array([1.356, 0.0, 1.726])
