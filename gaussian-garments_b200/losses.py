"""Fused photometric loss ("next" row N2 of SURVEY.md 8f) -- host side.

Mirrors the two calls the reference makes right after `render()`
(/root/reference/s2_registration.py:259-260, s3_appearance.py:132-133):

    loss_dict['img']  = l1_loss(image, gt_image, mask) * (1.0 - lambda_dssim)       utils/loss_utils.py:17-21
    loss_dict['ssim'] = 1.0 - ssim(image, gt_image, mask) * lambda_dssim            utils/loss_utils.py:36-69

    total, l1, ssim_value = photometric_loss(image, gt_image, mask, lambda_dssim)   # total == img + ssim terms

The arithmetic is csrc/photometric.cu behind gg_photometric_forward / gg_photometric_backward.  By default, unlike
the reference's ssim(), the rendered image and the ground truth are NOT multiplied by the mask in place;
`inplace_mask=True` reproduces that side effect (utils/loss_utils.py:44-46) for callers that rely on it.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _capi


def _prep(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"gaussian-garments_b200: `{name}` must be a CUDA tensor (there is no CPU path)")
    t = t.float() if t.dtype != torch.float32 else t
    return t.contiguous()


class _Photometric(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt, mask, lambda_dssim: float):
        lib = _capi.load()
        img = _prep(image, "image")
        u8 = gt.dtype == torch.uint8                       # 8-bit frame: value / 255, dequantised inside the kernels
        if u8 and not (lambda_dssim == 0.0 and (img.shape[1] * img.shape[2]) % 4 == 0 and gt.is_cuda):
            gt, u8 = gt.to(img.device).float() / 255.0, False          # SSIM path / odd sizes: plain conversion
        g = gt.contiguous() if u8 else _prep(gt, "gt")
        if img.dim() != 3 or img.shape[0] != 3 or g.shape != img.shape:
            raise RuntimeError("photometric_loss expects image and gt of shape [3,H,W]")
        m = None if mask is None else _prep(mask, "mask").reshape(-1)
        H, W = int(img.shape[1]), int(img.shape[2])
        if m is not None and m.numel() != H * W:
            raise RuntimeError("mask must have H*W elements ([1,H,W])")
        dev = img.device
        di = dev.index if dev.index is not None else torch.cuda.current_device()
        wb = C.c_size_t()
        _capi.check(lib.gg_photometric_workspace_bytes(W, H, C.byref(wb)), "gg_photometric_workspace_bytes")
        ws = torch.empty(wb.value if lambda_dssim != 0.0 else 1024, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            sp = torch.cuda.current_stream(dev).cuda_stream
            if u8:
                _capi.check(lib.gg_photometric_l1_u8(W, H, img.data_ptr(), g.data_ptr(), None if m is None else m.data_ptr(),
                                                     ws.data_ptr(), 0.0, None, None, di, sp), "gg_photometric_l1_u8")
            else:
                _capi.check(lib.gg_photometric_forward(W, H, img.data_ptr(), g.data_ptr(), None if m is None else m.data_ptr(),
                                                       ws.data_ptr(), 1 if lambda_dssim != 0.0 else 0, di, sp), "gg_photometric_forward")
            out3 = torch.empty(3, dtype=torch.float32, device=dev)
            _capi.check(lib.gg_photometric_reduce(W, H, ws.data_ptr(), float(lambda_dssim), out3.data_ptr(), di, sp),
                        "gg_photometric_reduce")
        total, l1, ssim_v = out3[0], out3[1], out3[2]
        ctx.save_for_backward(img, g, m if m is not None else torch.empty(0, device=dev), ws)
        ctx.lam, ctx.hw, ctx.u8 = float(lambda_dssim), (H, W), u8
        ctx.mark_non_differentiable(l1, ssim_v)
        return total, l1, ssim_v

    @staticmethod
    def backward(ctx, g_total, _g_l1, _g_ssim):
        lib = _capi.load()
        img, g, m, ws = ctx.saved_tensors
        H, W = ctx.hw
        dev = img.device
        di = dev.index if dev.index is not None else torch.cuda.current_device()
        n = 3.0 * H * W
        gs = g_total.reshape(1).float().contiguous()
        out = torch.empty_like(img)
        with torch.cuda.device(dev):
            sp = torch.cuda.current_stream(dev).cuda_stream
            if ctx.u8:
                _capi.check(lib.gg_photometric_l1_u8(W, H, img.data_ptr(), g.data_ptr(), m.data_ptr() if m.numel() else None,
                                                     ws.data_ptr(), (1.0 - ctx.lam) / n, gs.data_ptr(), out.data_ptr(), di, sp),
                            "gg_photometric_l1_u8")
            else:
                _capi.check(lib.gg_photometric_backward(W, H, img.data_ptr(), g.data_ptr(), m.data_ptr() if m.numel() else None,
                                                        ws.data_ptr(), (1.0 - ctx.lam) / n, -ctx.lam / n, gs.data_ptr(),
                                                        out.data_ptr(), di, sp), "gg_photometric_backward")
        return out, None, None, None


def photometric_loss(image: torch.Tensor, gt: torch.Tensor, mask: Optional[torch.Tensor] = None,
                     lambda_dssim: float = 0.2, inplace_mask: bool = False):
    """-> (total, l1, ssim): total = l1*(1-lambda) + 1 - ssim*lambda is differentiable w.r.t. `image`;
    l1 and ssim are the detached components (the values of the reference's l1_loss / ssim).

    inplace_mask=True follows the reference's call sequence to the letter: l1_loss on the unmasked tensors, then
    ssim() multiplies `image` and `gt` by the mask IN PLACE (utils/loss_utils.py:44-46) -- both tensors hold the masked
    values afterwards and the gradient flows through that in-place product (mask * mask for a soft mask's SSIM term,
    exactly as in the reference)."""
    lam = float(lambda_dssim)
    if not inplace_mask or mask is None:
        return _Photometric.apply(image, gt, mask, lam)
    # l1 + 1 on the tensors as handed in; the op keeps copies for its backward because both are overwritten below
    t_l1, l1, _ = _Photometric.apply(image.clone(), gt.clone(), mask, 0.0)
    m = mask.to(image.dtype)
    image.mul_(m)                                                        # autograd in-place op, as `img1 *= mask`
    with torch.no_grad():
        gt.mul_(m.to(gt.dtype))
    t_ss, _, ssim_v = _Photometric.apply(image, gt, None, 1.0)           # 1 - ssim, on the masked tensors
    total = (t_l1 - 1.0) * (1.0 - lam) + 1.0 - (1.0 - t_ss) * lam
    return total, l1, ssim_v
