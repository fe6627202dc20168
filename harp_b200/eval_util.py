"""Evaluation-pass image metrics on the device: drop-ins for `utils/eval_util.py`'s `fill_bg` (28-31),
`l1_diff` (34-39), `sil_iou` (41-49) and the part of `image_eval` (10-25) built on them.  LPIPS / MS-SSIM (51-60) need
pretrained networks and are out of scope like the VGG term."""
import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream


def fill_bg(img, mask):
    """eval_util.py:28-31: white background outside the mask."""
    img = img * mask.unsqueeze(-1)
    return img + (mask.unsqueeze(-1) - 1) * -1.0


def _metrics(ref_image, ref_mask, pred_image, pred_mask):
    f = lambda t: t.detach().float().contiguous()
    ref_image, ref_mask, pred_image, pred_mask = f(ref_image), f(ref_mask), f(pred_image), f(pred_mask)
    B, S = ref_mask.shape[0], ref_mask.shape[1]
    if ref_image.shape != (B, S, S, 3) or pred_image.shape != (B, S, S, 3) or pred_mask.shape != (B, S, S):
        raise ValueError("image metrics need (B,S,S,3) images and (B,S,S) masks of one size")
    l1 = torch.zeros(1, device=ref_image.device)
    iu = torch.zeros(2, B, dtype=torch.int32, device=ref_image.device)
    check(_lib.lib().harp_image_metrics(B, S, ptr(ref_image), ptr(pred_image), ptr(ref_mask), ptr(pred_mask), ptr(l1), _lib.iptr(iu[0]),
                                        _lib.iptr(iu[1]), stream()))
    return l1 / (3.0 * B * S * S), iu


def l1_diff(ref_image, ref_mask, pred_image, pred_mask):
    """eval_util.py:34-39: mean |ref - pred| over every pixel and channel (the masks are unused there too) -> numpy scalar."""
    l1, _ = _metrics(ref_image, ref_mask, pred_image, pred_mask)
    return l1[0].cpu().numpy()


def sil_iou(ref_masks, pred_masks):
    """eval_util.py:41-49: mean over frames of |ref & pred| / |ref | pred| at threshold 0.5 -> numpy scalar."""
    B, S = ref_masks.shape[0], ref_masks.shape[1]
    dummy = torch.zeros(B, S, S, 3, device=ref_masks.device)
    _, iu = _metrics(dummy, ref_masks, dummy, pred_masks)
    return torch.mean(iu[0] / iu[1]).cpu().numpy()


def image_eval(images_for_eval):
    """eval_util.py:10-25 without the two network-based metrics: lists of per-batch tensors -> {"Silhouette IoU", "L1"}."""
    v = {k: torch.vstack(list(t)) for k, t in images_for_eval.items()}
    l1, iu = _metrics(v["ref_image"], v["ref_mask"], v["pred_image"], v["pred_mask"])
    return {"Silhouette IoU": torch.mean(iu[0] / iu[1]).cpu().numpy(), "L1": l1[0].cpu().numpy()}
