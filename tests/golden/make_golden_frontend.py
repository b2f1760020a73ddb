"""tests/golden/frontend.npz: the first 24 000 rows of a shipped sample frame and what the UNMODIFIED reference
transform classes (VoxelSample(0.3,'first') -> DistanceSample(1,60) -> CoordinatesNormalization(60), after
BinReader's NaN-row drop) make of them, run in the build container.

    python tests/golden/make_golden_frontend.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_oracle_pin import REF, _reference_transforms  # noqa: E402

RT = _reference_transforms()
raw = np.fromfile(f"{REF}/data/sample/seq06/velodyne/000000.bin", dtype=np.float32).reshape(-1, 4)[:24000].copy()
raw[100] = np.nan  # BinReader drops NaN rows
xyz = raw[:, :3]
xyz = xyz[np.isnan(xyz).sum(1) == 0]
pcd = RT.PointCloud(xyz.copy())
for t in (RT.VoxelSample(0.3, "first"), RT.DistanceSample(1, 60), RT.CoordinatesNormalization(60)):
    pcd = t(pcd)
out = pcd.xyz.T.contiguous().numpy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "frontend.npz"), raw=raw, out=out)
print(raw.shape, out.shape)
