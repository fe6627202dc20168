from ..._core import _apply_lighting
