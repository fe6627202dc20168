"""Drop-in for `utils/config_utils.py:5-45`: the run configuration of `optimize_sequence.py` with the reference's keys and default
values, written to `<base_output_dir>/config.yaml` exactly as the reference does, plus the one piece of logic the loop derives from
it — which optimiser stage an epoch belongs to (`optimize_sequence.py:505-515`)."""
import os

import yaml

# key -> default, in the reference's order (utils/config_utils.py:6-32)
_DEFAULTS = (
    ("use_arm", True), ("opt_arm_pose", False), ("use_smooth_seq", True), ("average_cam_sequence", False),
    ("img_size", 448), ("focal_length", 2000.0),  # focal_length = 1000 * img_size / 224 (config_utils.py:11-12)
    ("model_type", "harp"), ("test_seq", False), ("known_appearance", False), ("load_siren", False), ("self_shadow", True),
    ("pose_already_opt", False), ("share_light_position", True), ("eval_mesh", False), ("use_vert_disp", True),
    ("total_epoch", 301), ("training_stage", [100, 100, 100]),  # [shape, shape and appearance, appearance only]
    ("metro_output_dir", "../data/sample_data/1/"), ("image_dir", "../data/sample_data/1/"),
    ("train_list", ["1", "2"]), ("val_list", ["1", "2"]), ("gt_mesh_dir", ""),
    ("base_output_dir", "exp/out_test/"), ("start_from", ""),
)
_TEMPLATES = {True: ("template/arm/arm_template.obj", "template/arm/uv_mask.png"),
              False: ("template/hand/textured_hand.obj", "template/hand/uv_mask.png")}


def get_config(yaml_file=None, write=True, **overrides):
    """-> the configuration dict.  `yaml_file` (ignored by the reference, which takes the argument but never reads it) is merged
    over the defaults here, then `overrides`; the template paths follow `use_arm` (config_utils.py:33-40).  With `write`, the
    result is dumped to `<base_output_dir>/config.yaml` like the reference (42-45)."""
    cfg = {k: (list(v) if isinstance(v, list) else v) for k, v in _DEFAULTS}
    if yaml_file:
        with open(yaml_file) as f:
            cfg.update(yaml.safe_load(f) or {})
    cfg.update(overrides)
    cfg["MANO_TEMPLATE"], cfg["uv_mask"] = _TEMPLATES[bool(cfg["use_arm"])]
    if write:
        os.makedirs(cfg["base_output_dir"], exist_ok=True)
        with open(os.path.join(cfg["base_output_dir"], "config.yaml"), "w") as f:
            yaml.dump(cfg, f)
    return cfg


def stage_of_epoch(configs, epoch_id):
    """'coarse' | 'both' | 'app' — COARSE_OPT / APP_OPT of optimize_sequence.py:505-515, as `HarpFitter.step(stage=...)` takes it."""
    a, b = configs["training_stage"][0], configs["training_stage"][1]
    return "coarse" if epoch_id < a else ("both" if epoch_id < a + b else "app")
