from ..._core import ShaderBase, SoftSilhouetteShader
