"""Golden vectors for the evaluation metrics: runs the REFERENCE'S OWN `utils/eval_util.py` functions (`sil_iou`, `l1_diff`,
`fill_bg`; 28-49) on small random inputs.  Build-container only.  `eval_util.py` imports lpips and pytorch_msssim at module
level and instantiates LPIPS on "cuda" (7-8); neither is installed here nor used by these three functions, so placeholder
modules are injected for the import.

    python tests/golden/make_golden_misc.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, REF)


class _Anything:
    def __init__(self, *a, **k):
        pass

    def to(self, *a, **k):
        return self


lp = types.ModuleType("lpips")
lp.LPIPS = _Anything
ms = types.ModuleType("pytorch_msssim")
ms.ssim = ms.ms_ssim = ms.SSIM = ms.MS_SSIM = _Anything
sys.modules.setdefault("lpips", lp)
sys.modules.setdefault("pytorch_msssim", ms)


def main():
    from utils import eval_util  # the reference module, unmodified
    g = torch.Generator().manual_seed(11)
    B, S = 3, 32
    ref_image, pred_image = torch.rand(B, S, S, 3, generator=g), torch.rand(B, S, S, 3, generator=g)
    ref_mask = (torch.rand(B, S, S, generator=g) > 0.4).float() * torch.rand(B, S, S, generator=g)
    pred_mask = (torch.rand(B, S, S, generator=g) > 0.5).float() * torch.rand(B, S, S, generator=g)
    out = dict(ref_image=ref_image.numpy(), pred_image=pred_image.numpy(), ref_mask=ref_mask.numpy(), pred_mask=pred_mask.numpy(),
               sil_iou=np.float64(eval_util.sil_iou(ref_mask, pred_mask)),
               l1_diff=np.float64(eval_util.l1_diff(ref_image, ref_mask, pred_image, pred_mask)),
               fill_bg=eval_util.fill_bg(ref_image, ref_mask).numpy())
    np.savez_compressed(os.path.join(HERE, "eval_metrics.npz"), **out)
    print("sil_iou", out["sil_iou"], "l1", out["l1_diff"])


if __name__ == "__main__":
    main()
