def load_obj(*a, **k):  # utils/hand_model_utils.py:5 only; asset loading is outside the hot path
    raise NotImplementedError("oracle shim: use harp_b200.assets instead")
