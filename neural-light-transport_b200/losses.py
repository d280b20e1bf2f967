"""Losses with the reference's class names (nlt/losses.py).

Only L2 is on the accelerated hot path (it is the loss used
when timing the path); it is fused with its own gradient in one CUDA pass.
Barron / LPIPS are the "next" row N1 and raise NotImplementedError.
"""
import torch

import nlt_native as nat


class L2():
    """nlt/losses.py:39-53 -- mean squared error; keep_batch -> (N,)."""

    def __init__(self):
        self._ws = None
        self.d_pred = None      # gradient of sum_b(loss_b) * grad_scale w.r.t. pred
        self.grad_scale = None  # set by the train step: 1 / global_bs

    def __call__(self, gt, pred, keep_batch=False, weights=None):
        if weights is not None:
            raise NotImplementedError('sample weights')
        lib = nat.lib()
        B = pred.shape[0]
        per = pred.numel() // B
        need = lib.nlt_l2_loss_workspace_bytes(B, per)
        if self._ws is None or self._ws.numel() * 4 < need or self._ws.device != pred.device:
            self._ws = torch.empty((need + 3) // 4, dtype=torch.float32, device=pred.device)
        loss = torch.empty(B, dtype=torch.float32, device=pred.device)
        want_grad = self.grad_scale is not None
        self.d_pred = torch.empty_like(pred) if want_grad else None
        scale = float(self.grad_scale) if want_grad else 0.0
        if not keep_batch and want_grad:
            scale = scale / B   # mean over the batch as well
        nat.check(lib.nlt_l2_loss(nat.ptr(pred), nat.ptr(gt), B, per, scale, nat.ptr(loss),
                                  nat.ptr(self.d_pred), nat.ptr(self._ws), nat.stream()))
        if keep_batch:
            return loss
        return loss.mean()


class L1():
    def __call__(self, gt, pred, weights=None):
        raise NotImplementedError('l1 is not on the accelerated path')


class SSIM():
    def __init__(self, dynamic_range):
        raise NotImplementedError('ssim is not on the accelerated path')


class Barron():
    def __init__(self, imw, imh):
        raise NotImplementedError('barron: next row N1 (SURVEY.md 8f)')


class LPIPS():
    def __init__(self, per_ch=False):
        raise NotImplementedError('lpips: next row N1 (SURVEY.md 8f)')
