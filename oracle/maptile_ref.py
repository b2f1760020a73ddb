"""oracle/maptile_ref.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the scan-to-map input stage of the reference:
  PoseGraph.__global_mapping(full_pcd=False)       /root/reference/system/modules/pose_graph.py:373-409
  centring in PoseGraph.global_map_query_graph     /root/reference/system/modules/pose_graph.py:499-511

Pinned by tests/test_oracle_pin.py::test_map_tile_matches_reference_posegraph (the reference's PoseGraph class
itself, run in the build container with stub `readerwriterlock` / `open3d` / `matplotlib` modules).
"""
import torch


def map_tile(key_points, poses, center=None):
    """key_points: list of (Cd, S) [fea ; xyz]; poses: list of (4,4) SE3_pred; center (4,4) -> (Cd, m*S)"""
    tiles = []
    for kp, T in zip(key_points, poses):
        pts = kp.clone()                                              # pose_graph.py:390
        pts[-3:, :] = T[:3, :3] @ pts[-3:, :] + T[:3, 3:]             # pose_graph.py:391
        tiles.append(pts)
    tile = torch.concat(tiles, dim=1)                                 # pose_graph.py:407
    if center is not None:
        R, t = center[:3, :3], center[:3, 3:]                         # pose_graph.py:505
        tile[-3:, :] = R.T @ (tile[-3:, :] - t)                       # pose_graph.py:507
    return tile
