"""pytorch3d.ops subset used by DeepPointMap: sample_farthest_points, knn_points, knn_gather,
ball_query (signatures of pytorch3d 0.7.4; CUDA tensors only, no CPU fallback)."""
from deeppointmap_b200.ops import ball_query, knn_gather, knn_points, sample_farthest_points  # noqa: F401

__all__ = ["ball_query", "knn_gather", "knn_points", "sample_farthest_points"]
