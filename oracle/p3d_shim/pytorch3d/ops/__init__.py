from .._core import interpolate_face_attributes, SubdivideMeshes
