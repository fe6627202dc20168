from .._core import Meshes
