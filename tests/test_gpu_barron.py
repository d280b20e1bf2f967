"""The CUDA Barron-loss entry point nlt_barron_loss against the reference-pinned oracle (oracle/barron_oracle.py).
The arithmetic shared with these kernels is also checked on the CPU (tests/test_barron_core.py)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import barron_oracle as B

pytestmark = [pytest.mark.gpu]

LOG_Z = 1.185495232349193


@pytest.mark.parametrize('Bn,H,W,use_alpha', [(2, 32, 48, False), (1, 83, 71, True), (2, 33, 17, True), (8, 256, 256, False)])
def test_barron_loss_and_gradient(Bn, H, W, use_alpha):
    import nlt_native as nat
    lib = nat.lib()
    dev = torch.device('cuda')
    g = torch.Generator().manual_seed(H + W)
    gt = torch.rand(Bn, H, W, 3, generator=g, dtype=torch.float64)
    pred = (gt + 0.03 * torch.randn(Bn, H, W, 3, generator=g, dtype=torch.float64)).requires_grad_(True)
    alpha = torch.rand(Bn, H, W, 1, generator=g, dtype=torch.float64) if use_alpha else None
    loss_scale = 0.125
    per = B.barron_loss(gt, pred, LOG_Z, keep_batch=True, weights=alpha)
    (per.sum() * loss_scale).backward()
    p32, g32 = pred.detach().float().to(dev).contiguous(), gt.float().to(dev).contiguous()
    a32 = alpha.float().to(dev).contiguous() if use_alpha else None
    need = lib.nlt_barron_loss_workspace_bytes(Bn, H, W, 5)
    assert need > 0
    ws = torch.empty((need + 3) // 4, dtype=torch.float32, device=dev)
    loss = torch.empty(Bn, dtype=torch.float32, device=dev)
    d_pred = torch.empty_like(p32)
    nat.check(lib.nlt_barron_loss(nat.ptr(p32), nat.ptr(g32), nat.ptr(a32), Bn, H, W, 5, 0.01, LOG_Z, loss_scale,
                                  nat.ptr(loss), nat.ptr(d_pred), nat.ptr(ws), nat.stream()))
    torch.cuda.synchronize()
    np.testing.assert_allclose(loss.cpu().numpy(), per.detach().numpy(), rtol=5e-5, atol=5e-5)
    want = pred.grad
    err = float((d_pred.cpu().double() - want).norm() / want.norm())
    assert err <= 5e-5, err
    # loss only (d_pred = NULL) gives the same values
    loss2 = torch.empty_like(loss)
    nat.check(lib.nlt_barron_loss(nat.ptr(p32), nat.ptr(g32), nat.ptr(a32), Bn, H, W, 5, 0.01, LOG_Z, 0.0,
                                  nat.ptr(loss2), None, nat.ptr(ws), nat.stream()))
    torch.testing.assert_close(loss2, loss, rtol=1e-6, atol=1e-6)


def test_barron_bad_arguments():
    import nlt_native as nat
    lib = nat.lib()
    assert lib.nlt_barron_loss_workspace_bytes(1, 8, 8, 5) < 0          # 8 x 8 cannot hold 5 levels
    assert lib.nlt_barron_loss_workspace_bytes(1, 32, 32, 5) > 0
