from ..._core import Fragments, MeshRasterizer, RasterizationSettings
