from .._core import (look_at_rotation, PerspectiveCameras, PointLights, Materials, RasterizationSettings, MeshRenderer,
                     MeshRasterizer, SoftSilhouetteShader, TexturesVertex, TexturesUV, BlendParams, TensorProperties)
