from .._core import Device
