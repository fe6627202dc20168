from .._core import TensorProperties, convert_to_tensors_and_broadcast
