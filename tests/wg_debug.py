"""Debug probe for the tensor-core wgrad kernel (GPU box)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'neural-light-transport_b200')]
import engine  # noqa: E402
import nlt_native as nat  # noqa: E402

dev = torch.device('cuda')
lib = nat.lib()
torch.set_printoptions(linewidth=200, precision=1, sci_mode=False)


def run(kind, k, s, H, W, segc, cout, mode, N=2):
    L = engine.ConvLayer(kind, k, s, cout, None)
    L.build(sum(segc), dev, torch.Generator().manual_seed(1))
    xs = []
    for c in segc:
        if mode == 'ones':
            xs.append(torch.ones(N, H, W, c, device=dev))
        elif mode == 'chan':
            xs.append((torch.arange(c, device=dev, dtype=torch.float32) + 1).expand(N, H, W, c).contiguous())
        else:
            xs.append(torch.randn(N, H, W, c, device=dev))
    segs = [engine.Seg(engine.Act(x)) for x in xs]
    d = L._fwd_desc(segs, N, H, W)
    if mode == 'gcol':
        dz = (torch.arange(cout, device=dev, dtype=torch.float32) + 1).expand(N, d.Hout, d.Wout, cout).contiguous()
    elif mode == 'rand':
        dz = torch.randn(N, d.Hout, d.Wout, cout, device=dev)
    else:
        dz = torch.ones(N, d.Hout, d.Wout, cout, device=dev)
    need = lib.nlt_gconv_wgrad_workspace_bytes(C.byref(d))
    ws = torch.full(((need + 3) // 4,), 7.0, device=dev)
    t0 = nat.tc_launch_count()
    nat.check(lib.nlt_gconv_wgrad(C.byref(d), nat.ptr(dz), nat.ptr(L.gkernel), nat.ptr(L.gbias), 0, nat.ptr(ws),
                                  ws.numel() * 4, nat.stream()))
    torch.cuda.synchronize()
    used = nat.tc_launch_count() - t0
    # reference via torch autograd (fp64)
    from oracle import nlt_oracle as O
    w64 = L.kernel.double().cpu().requires_grad_(True)
    b64 = L.bias.double().cpu().requires_grad_(True)
    fn = O.conv2d_same if kind == 'conv' else O.conv2d_transpose_same
    y = fn(torch.cat([x.double().cpu() for x in xs], 3), w64, b64, s)
    (y * dz.double().cpu()).sum().backward()
    gk, gb = L.gkernel.double().cpu(), L.gbias.double().cpu()
    rk = float((gk - w64.grad).norm() / w64.grad.norm())
    rb = float((gb - b64.grad).norm() / b64.grad.norm())
    print(kind, k, s, H, W, segc, cout, mode, 'tc=%d' % used, 'relK %.2e relB %.2e' % (rk, rb), flush=True)
    if rk > 1e-4 and mode != 'rand':
        print(' got  [tap0, :, :4]', gk.reshape(k * k, -1, cout)[0, :, :4].T if kind == 'conv' else gk.reshape(k * k, cout, -1)[0, :4, :])
        print(' want [tap0, :, :4]', w64.grad.reshape(k * k, -1, cout)[0, :, :4].T if kind == 'conv' else w64.grad.reshape(k * k, cout, -1)[0, :4, :])
        print(' gb got', gb[:8], 'want', b64.grad[:8])


for mode in ('ones', 'chan', 'gcol', 'rand'):
    run('conv', 1, 1, 16, 16, [32], 32, mode)
for mode in ('ones', 'rand'):
    run('conv', 1, 1, 16, 16, [16], 16, mode)
    run('conv', 2, 2, 32, 32, [32, 32], 64, mode)
    run('conv', 2, 1, 16, 16, [64], 64, mode)
    run('deconv', 2, 2, 16, 16, [64, 64], 32, mode)
    run('conv', 2, 2, 32, 32, [16, 16], 16, mode)
