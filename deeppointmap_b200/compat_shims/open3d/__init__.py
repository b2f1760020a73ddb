"""Shim for `open3d` 0.16 (absent in this image): the reference imports it at module scope in
dataloader/transforms.py:9, system/modules/utils.py:7, system/modules/recoder.py:14 and pose_graph.py, but inference
through `pipeline/infer.py` only needs it for
  * the final map dump (recoder.py:176-193): `geometry.PointCloud` + `voxel_down_sample` + `io.write_point_cloud`
    -> provided here with numpy (mean of the points per voxel, ASCII .pcd);
  * `estimate_normals` of `LowPassFilter` (transforms.py:269-272) -> provided with numpy / scipy (PCA of the
    neighbours inside the radius, as open3d's KDTreeSearchParamRadius path does);
  * the pose-graph optimiser (pose_graph.py:573-607, loop closure only) -> NOT provided: raises, run with
    `enable_global_optimization: false`.
Everything else raises AttributeError on use."""
import numpy as np

__version__ = "0.16.0-shim"


class _Vec(np.ndarray):
    pass


def _as_vec(a, cols):
    a = np.asarray(a, dtype=np.float64).reshape(-1, cols)
    return a.view(_Vec)


class _Utility:
    @staticmethod
    def Vector3dVector(a):
        return _as_vec(a, 3)


utility = _Utility()


class _KDTreeSearchParamRadius:
    def __init__(self, radius):
        self.radius = float(radius)


class _KDTreeSearchParamKNN:
    def __init__(self, knn=30):
        self.knn = int(knn)


class _KDTreeSearchParamHybrid:
    def __init__(self, radius, max_nn):
        self.radius, self.max_nn = float(radius), int(max_nn)


class PointCloud:
    def __init__(self):
        self.points = np.zeros((0, 3))
        self.normals = np.zeros((0, 3))

    def voxel_down_sample(self, voxel_size):
        pts = np.asarray(self.points, dtype=np.float64).reshape(-1, 3)
        out = PointCloud()
        if len(pts) == 0:
            return out
        key = np.floor((pts - pts.min(0)) / float(voxel_size)).astype(np.int64)
        _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
        acc = np.zeros((len(cnt), 3))
        np.add.at(acc, inv.reshape(-1), pts)
        out.points = _as_vec(acc / cnt[:, None], 3)
        return out

    def estimate_normals(self, search_param=None, fast_normal_computation=True):
        """per point: eigenvector of the smallest eigenvalue of the covariance of its neighbours (open3d
        geometry/EstimateNormals.cpp); a point with fewer than 3 neighbours gets (0, 0, 1) like open3d"""
        from scipy.spatial import cKDTree
        pts = np.asarray(self.points, dtype=np.float64).reshape(-1, 3)
        tree = cKDTree(pts)
        normals = np.tile(np.array([0.0, 0.0, 1.0]), (len(pts), 1))
        if isinstance(search_param, _KDTreeSearchParamKNN):
            nbrs = tree.query(pts, k=min(search_param.knn, len(pts)))[1]
            nbrs = [row for row in np.atleast_2d(nbrs)]
        elif isinstance(search_param, _KDTreeSearchParamHybrid):
            nbrs = [tree.query_ball_point(p, search_param.radius)[:search_param.max_nn] for p in pts]
        else:
            r = search_param.radius if search_param is not None else 0.1
            nbrs = tree.query_ball_point(pts, r)
        for i, nb in enumerate(nbrs):
            if len(nb) < 3:
                continue
            q = pts[np.asarray(nb)]
            w, v = np.linalg.eigh(np.cov(q.T, bias=True))
            normals[i] = v[:, 0]
        self.normals = _as_vec(normals, 3)
        return True

    def remove_statistical_outlier(self, nb_neighbors, std_ratio):
        from scipy.spatial import cKDTree
        pts = np.asarray(self.points, dtype=np.float64).reshape(-1, 3)
        d = cKDTree(pts).query(pts, k=min(int(nb_neighbors), len(pts)))[0]
        avg = d.mean(1)
        keep = avg < avg.mean() + float(std_ratio) * avg.std()
        out = PointCloud()
        out.points = _as_vec(pts[keep], 3)
        return out, np.nonzero(keep)[0].tolist()


class _Geometry:
    PointCloud = PointCloud
    KDTreeSearchParamRadius = _KDTreeSearchParamRadius
    KDTreeSearchParamKNN = _KDTreeSearchParamKNN
    KDTreeSearchParamHybrid = _KDTreeSearchParamHybrid


geometry = _Geometry()


class _IO:
    @staticmethod
    def write_point_cloud(filename, pointcloud, write_ascii=True, **k):
        pts = np.asarray(pointcloud.points, dtype=np.float32).reshape(-1, 3)
        with open(filename, "w") as f:
            f.write("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\n"
                    f"COUNT 1 1 1\nWIDTH {len(pts)}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {len(pts)}\nDATA ascii\n")
            np.savetxt(f, pts, fmt="%.6f")
        return True

    @staticmethod
    def read_point_cloud(filename, **k):
        raise NotImplementedError("open3d shim: .pcd input is not supported (use the .bin reader)")


io = _IO()


class _Missing:
    def __init__(self, what):
        self._what = what

    def __getattr__(self, name):
        raise NotImplementedError(f"open3d shim: open3d.{self._what}.{name} needs the real open3d package "
                                  "(pose-graph optimisation: run with enable_global_optimization: false)")


class _Registration:
    """only what inference WITHOUT loop closure touches; the pose-graph optimiser raises"""

    @staticmethod
    def get_information_matrix_from_point_clouds(source, target, max_correspondence_distance, transformation):
        """open3d pipelines/registration/Registration.cpp GetInformationMatrixFromPointClouds: for every source point
        whose transformed position has a target point within the distance, that target point (x, y, z) adds G^T G of
        G = [[0, z, -y, 1, 0, 0], [-z, 0, x, 0, 1, 0], [y, -x, 0, 0, 0, 1]] -- the formula the reference restates for
        its pytorch3d branch (system/modules/utils.py:70-101)."""
        from scipy.spatial import cKDTree
        T = np.asarray(transformation, dtype=np.float64).reshape(4, 4)
        src = np.asarray(source.points, dtype=np.float64).reshape(-1, 3) @ T[:3, :3].T + T[:3, 3]
        tgt = np.asarray(target.points, dtype=np.float64).reshape(-1, 3)
        d, idx = cKDTree(tgt).query(src, k=1, distance_upper_bound=float(max_correspondence_distance))
        t = tgt[idx[np.isfinite(d)]]
        x, y, z = t[:, 0], t[:, 1], t[:, 2]
        o, l = np.zeros_like(x), np.ones_like(x)
        GTG = np.zeros((6, 6))
        for row in (np.stack([o, z, -y, l, o, o], 1), np.stack([-z, o, x, o, l, o], 1), np.stack([y, -x, o, o, o, l], 1)):
            GTG += row.T @ row
        return GTG

    def __getattr__(self, name):
        raise NotImplementedError(f"open3d shim: open3d.pipelines.registration.{name} needs the real open3d package "
                                  "(pose-graph optimisation: run with enable_global_optimization: false)")


class _Pipelines:
    registration = _Registration()


pipelines = _Pipelines()

import sys as _sys
open3d = _sys.modules.get(__name__)  # the reference spells `o3d.open3d.utility.Vector3dVector` (recoder.py:179)
