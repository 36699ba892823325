"""Point-cloud meshes (host-side set-up).  API of src/pnmol/mesh.py for 1-D boxes."""
from functools import cached_property

import numpy as np
import scipy.spatial


class RectangularMesh:
    def __init__(self, points, bbox=None):
        self.points = np.asarray(points, dtype=np.float64)
        if bbox is None:  # mesh.py:178-184
            bbox = np.stack([self.points.min(axis=0), self.points.max(axis=0)], axis=1)
        self.bbox = np.asarray(bbox, dtype=np.float64)
        self._tree = scipy.spatial.KDTree(self.points)

    @classmethod
    def from_bbox_1d(cls, bbox, step=None, num=None):
        """mesh.py:86-98 (``step`` keeps the reference's floating-point floor: 1/99 gives 99 points)."""
        bbox = np.asarray(bbox, dtype=np.float64)
        if int(step is None) + int(num is None) != 1:
            raise ValueError("Provide exactly one of step or num.")
        if step is not None:
            num = int((bbox[1] - bbox[0]) / step) + 1
        return cls(np.linspace(bbox[0], bbox[1], num, endpoint=True).reshape(-1, 1))

    def neighbours(self, point, num):
        if num <= 0:
            raise ValueError("num >= 1 required!")
        _, idx = self._tree.query(np.asarray(point), k=num)
        idx = np.asarray(idx)
        if num == 1:
            idx = idx[..., None]
        return self.points[idx], idx

    def _on_boundary(self):
        on = np.zeros(len(self.points), dtype=bool)
        for dim in range(self.points.shape[1]):
            on |= (self.points[:, dim] == self.bbox[dim, 0]) | (self.points[:, dim] == self.bbox[dim, 1])
        return on

    @cached_property
    def boundary(self):
        on = self._on_boundary()
        return self.points[on], on, np.nonzero(on)[0]

    @cached_property
    def interior(self):
        off = ~self._on_boundary()
        return self.points[off], off, np.nonzero(off)[0]

    @cached_property
    def boundary_projection_matrix(self):
        return np.eye(len(self.points))[self.boundary[1], :]

    def __len__(self):
        return len(self.points)

    def __getitem__(self, key):
        return self.points[key]

    @property
    def shape(self):
        return self.points.shape

    @property
    def dimension(self):
        return self.points.shape[-1]
