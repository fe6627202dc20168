"""Evaluation-pass image metrics (utils/eval_util.py:28-49) and the resume-time parameter re-initialisation
(optimize_sequence.py:366-389)."""
import numpy as np
import pytest
import torch

import harp_oracle as O
from util import load_golden


def test_oracle_metrics_match_reference_golden():
    g = load_golden("eval_metrics")
    T = lambda k: torch.tensor(g[k])
    assert abs(float(O.sil_iou(T("ref_mask"), T("pred_mask"))) - float(g["sil_iou"])) < 1e-7
    assert abs(float(O.l1_diff(T("ref_image"), T("ref_mask"), T("pred_image"), T("pred_mask"))) - float(g["l1_diff"])) < 1e-7
    assert torch.equal(O.fill_bg(T("ref_image"), T("ref_mask")), T("fill_bg"))


def test_resume_params_host_logic_matches_oracle():
    """CPU tensors: the product's vectorised re-interpolation equals the reference's double loop bit for bit."""
    from harp_b200.file_utils import prepare_resumed_params
    gen = torch.Generator().manual_seed(2)
    p = dict(pose=torch.randn(97, 45, generator=gen), trans=torch.randn(97, 3, generator=gen), rot=torch.randn(97, 3, generator=gen),
             shape=torch.randn(10, generator=gen), texture=torch.rand(1, 8, 8, 3, generator=gen))
    mine, ref = prepare_resumed_params(p), O.resume_params(p)
    assert set(ref) <= set(mine)
    for k, v in ref.items():
        assert torch.equal(mine[k], v), k
    assert mine["wrist_pose"].shape == (97, 3) and float(mine["amb_ratio"]) == pytest.approx(0.4)
    assert mine["normal_map"].shape == (1, 512, 512, 3)
    assert not torch.equal(mine["pose"][:60], p["pose"][:60]) and torch.equal(mine["pose"][60:], p["pose"][60:])


@pytest.mark.gpu
def test_device_metrics_match_reference_golden():
    from harp_b200 import eval_util
    g = load_golden("eval_metrics")
    C = lambda k: torch.tensor(g[k]).cuda()
    assert abs(float(eval_util.sil_iou(C("ref_mask"), C("pred_mask"))) - float(g["sil_iou"])) < 1e-6
    assert abs(float(eval_util.l1_diff(C("ref_image"), C("ref_mask"), C("pred_image"), C("pred_mask"))) - float(g["l1_diff"])) < 1e-6
    assert torch.equal(eval_util.fill_bg(C("ref_image"), C("ref_mask")).cpu(), torch.tensor(g["fill_bg"]))
    st = eval_util.image_eval(dict(ref_image=[C("ref_image")[:2], C("ref_image")[2:]], pred_image=[C("pred_image")[:2], C("pred_image")[2:]],
                                   ref_mask=[C("ref_mask")[:2], C("ref_mask")[2:]], pred_mask=[C("pred_mask")[:2], C("pred_mask")[2:]]))
    assert abs(float(st["Silhouette IoU"]) - float(g["sil_iou"])) < 1e-6 and abs(float(st["L1"]) - float(g["l1_diff"])) < 1e-6
    # full size: 64 frames at 512^2 against torch on the device
    gen = torch.Generator(device="cuda").manual_seed(0)
    a, b = torch.rand(64, 512, 512, 3, device="cuda", generator=gen), torch.rand(64, 512, 512, 3, device="cuda", generator=gen)
    ma, mb = torch.rand(64, 512, 512, device="cuda", generator=gen), torch.rand(64, 512, 512, device="cuda", generator=gen)
    assert abs(float(eval_util.l1_diff(a, ma, b, mb)) - float((a - b).abs().double().mean())) < 1e-6
    assert abs(float(eval_util.sil_iou(ma, mb)) - float(O.sil_iou(ma.cpu(), mb.cpu()))) < 1e-6
