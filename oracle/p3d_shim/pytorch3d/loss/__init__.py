from .._core import mesh_laplacian_smoothing, mesh_normal_consistency
