from .._core import BlendParams, softmax_rgb_blend, sigmoid_alpha_blend
