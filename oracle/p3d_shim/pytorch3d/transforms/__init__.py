from .._core import RotateAxisAngle, Transform3d
