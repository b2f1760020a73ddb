"""tests/golden/infomat.npz: outputs of the UNMODIFIED reference function
calculate_information_matrix_from_pcd (/root/reference/system/modules/utils.py:60-104), pytorch3d branch,
run on CPU in the build container with a plain-torch knn_points (direct differences, K=1).

    python tests/golden/make_golden_infomat.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ref_infomat import reference_information_matrix, cases  # noqa: E402

out = {}
for name, (p1, p2, T, radius) in cases().items():
    out[name + "_p1"], out[name + "_p2"], out[name + "_T"] = p1.numpy(), p2.numpy(), T.numpy()
    out[name + "_radius"] = np.float32(radius)
    out[name + "_info"] = reference_information_matrix(p1, p2, T, radius).numpy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "infomat.npz"), **out)
print({k: v.shape for k, v in out.items()})
