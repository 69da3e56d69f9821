"""`Coordinates` value type exchanged with TrackerLite -- API-compatible with
CellTracker/coord_image_transformer.py:29-141 (raw / real / interp views of (n,3) cell centres).

Only the coordinate container is on the hot path's boundary; the label-image bookkeeping of
CoordsToImageTransformer is host glue outside SURVEY section 8.
"""
import numpy as np


class Coordinates:
    def __init__(self, coords, interpolation_factor, voxel_size, dtype="raw"):
        self.interpolation_factor = interpolation_factor
        self.voxel_size = np.asarray(voxel_size)
        coords = np.asarray(coords).astype(np.float32)
        if dtype == "raw":
            self._raw = coords
        elif dtype == "real":
            self._raw = (coords * (1.0 / self.voxel_size)[None, :]).astype(np.float32)
        elif dtype == "interp":
            self._raw = (coords * np.asarray((1, 1, 1 / interpolation_factor))[None, :]).astype(np.float32)
        else:
            raise ValueError(f"unknown coordinate type {dtype!r}")

    def __add__(self, other):
        return Coordinates(self._raw + other._raw, self.interpolation_factor, self.voxel_size, "raw")

    def __sub__(self, other):
        return Coordinates(self._raw - other._raw, self.interpolation_factor, self.voxel_size, "raw")

    @property
    def real(self):
        return self._raw * self.voxel_size[None, :]

    @property
    def interp(self):
        return np.round(self._raw * np.asarray((1, 1, self.interpolation_factor))[None, :]).astype(np.int32)

    @property
    def raw(self):
        return np.round(self._raw).astype(np.int32)

    @property
    def cell_num(self):
        return self._raw.shape[0]
