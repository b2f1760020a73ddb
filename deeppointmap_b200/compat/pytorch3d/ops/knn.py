"""pytorch3d.ops.knn (imported as `from pytorch3d.ops.knn import knn_points`,
system/modules/utils.py:10)."""
from deeppointmap_b200.ops import knn_gather, knn_points  # noqa: F401
