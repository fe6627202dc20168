"""ORACLE / TEST ONLY — minimal restatement of the `smplx` package modules that `hand_models_harp/body_models.py:29-41`
imports relatively (`.lbs`, `.vertex_ids`, `.utils`, `.vertex_joint_selector`).  The reference README (58-59) tells users to
drop `body_models.py` into a downloaded smplx checkout; that checkout is not vendored, so these four modules restate its
published code.  PARITY UNPINNED at this layer.  `load_reference_body_models()` then imports
the reference's own, unmodified `body_models.py` into this package so that `SMPLXARM.forward` itself is the oracle above it.
"""
import importlib.util
import os
import sys


def load_reference_body_models(reference_root="/root/reference"):
    name = __name__ + ".body_models"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(reference_root, "hand_models_harp", "body_models.py")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = __name__
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
