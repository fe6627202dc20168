"""`harp_b200.config_utils` against the reference's `utils/config_utils.py` (imported live when /root/reference exists; the key
set and defaults are also pinned here so that the test means something on a box without the reference)."""
import os
import sys

import yaml

from harp_b200 import config_utils

REF = "/root/reference"
PINNED = {"use_arm": True, "opt_arm_pose": False, "use_smooth_seq": True, "average_cam_sequence": False, "img_size": 448,
          "focal_length": 2000.0, "model_type": "harp", "self_shadow": True, "share_light_position": True, "use_vert_disp": True,
          "total_epoch": 301, "training_stage": [100, 100, 100], "MANO_TEMPLATE": "template/arm/arm_template.obj",
          "uv_mask": "template/arm/uv_mask.png"}


def test_defaults_and_written_yaml(tmp_path):
    out = str(tmp_path / "exp")
    cfg = config_utils.get_config(base_output_dir=out)
    for k, v in PINNED.items():
        assert cfg[k] == v, k
    assert cfg["focal_length"] == 1000.0 * cfg["img_size"] / 224.0
    assert yaml.safe_load(open(os.path.join(out, "config.yaml"))) == cfg
    hand = config_utils.get_config(write=False, use_arm=False, img_size=512, focal_length=1000.0 * 512 / 224)
    assert hand["MANO_TEMPLATE"] == "template/hand/textured_hand.obj" and hand["uv_mask"] == "template/hand/uv_mask.png"


def test_matches_the_reference_when_present(tmp_path, monkeypatch):
    if not os.path.exists(os.path.join(REF, "utils", "config_utils.py")):
        import pytest
        pytest.skip("reference tree not on this box")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_config_utils", os.path.join(REF, "utils", "config_utils.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    monkeypatch.chdir(tmp_path)  # the reference writes exp/out_test/config.yaml relative to the working directory
    want = ref.get_config()
    got = config_utils.get_config(write=False)
    assert got == want and list(got) == list(want)  # same keys, values and order


def test_stage_schedule():
    cfg = config_utils.get_config(write=False)
    stages = [config_utils.stage_of_epoch(cfg, e) for e in range(cfg["total_epoch"])]
    assert stages[0] == "coarse" and stages[99] == "coarse" and stages[100] == "both" and stages[199] == "both"
    assert stages[200] == "app" and stages[300] == "app"  # optimize_sequence.py:507-515
