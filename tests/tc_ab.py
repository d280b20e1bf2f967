"""A/B accuracy probe (run on the GPU box): tensor-core path vs the fp32-FMA kernels of the same library on
identical inputs.  Prints relative Frobenius errors; used to localise precision problems."""
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


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def run(kind, k, s, H, W, segc, cout, N=2):
    torch.manual_seed(1)
    xs = [torch.randn(N, H, W, c, device=dev) for c in segc]
    L = engine.ConvLayer(kind, k, s, cout, 'leakyrelu')
    L.build(sum(segc), dev, torch.Generator().manual_seed(1))
    L.bias.copy_(torch.randn(cout, device=dev) * 0.1)
    acts = [engine.Act(x, act='leakyrelu', needs_grad=True) for x in xs]
    segs = [engine.Seg(a) for a in acts]
    t0 = nat.tc_launch_count()
    tape = engine.Tape()
    y = L.forward(segs, tape)
    used_fwd = nat.tc_launch_count() - t0
    d = L._fwd_desc(segs, N, H, W)
    y_fp = torch.empty_like(y.t)
    nat.check(lib.nlt_gconv_fwd(C.byref(d), nat.ptr(L.bias), 2, 0.0, None, 0, nat.ptr(y_fp), nat.stream()))
    out = ['fwd tc=%d err=%.2e' % (used_fwd, rel(y.t, y_fp))]
    # dgrad with beta + mask, and wgrad
    dz = torch.randn_like(y.t)
    y.grad = dz.clone()
    for a in acts:
        a.n_cons = 2            # so that the first contribution has no mask, then add a second one with mask
    t0 = nat.tc_launch_count()
    tape.backward()
    used_bwd = nat.tc_launch_count() - t0
    gk_tc, gb_tc = L.gkernel.clone(), L.gbias.clone()
    grads_tc = [a.grad.clone() for a in acts]
    # fp32 reference of the same calls
    os.environ['X'] = '1'
    coff = 0
    for a, sg, g_tc in zip(acts, segs, grads_tc):
        dd = L._dgrad_desc(dz, N, H, W, coff, sg.C)
        ref = torch.empty_like(a.t)
        nat.check(lib.nlt_gconv_fwd(C.byref(dd), None, 0, 0.0, None, 0, nat.ptr(ref), nat.stream()))
        out.append('dgrad[%d] err=%.2e' % (sg.C, rel(g_tc, ref)))
        # second contribution: beta=1 + mask on the tensor-core path vs fp32 path
        acc_tc, acc_fp = g_tc.clone(), ref.clone()
        engine.gconv_fwd(dd, None, 0, 1.0, a.t, 2, acc_tc)
        nat.check(lib.nlt_gconv_fwd(C.byref(dd), None, 0, 1.0, nat.ptr(a.t), 2, nat.ptr(acc_fp), nat.stream()))
        out.append('rmw err=%.2e' % rel(acc_tc, acc_fp))
        coff += sg.C
    out.append('bwd tc=%d' % used_bwd)
    return out, (gk_tc, gb_tc, L, segs, dz, N, H, W)


SHAPES = [
    ('conv', 2, 2, 32, 32, [32, 32], 64),
    ('conv', 2, 1, 16, 16, [16], 16),
    ('conv', 2, 2, 32, 32, [16, 16], 16),
    ('conv', 2, 1, 32, 32, [32], 32),
    ('deconv', 2, 2, 16, 16, [16, 32, 32], 8),
    ('deconv', 2, 1, 32, 16, [16], 16),
    ('deconv', 2, 2, 16, 16, [64, 64], 32),
    ('conv', 2, 2, 64, 64, [32, 16], 32),
]
for sh in SHAPES:
    res, extra = run(*sh)
    gk_tc, gb_tc, L, segs, dz, N, H, W = extra
    # wgrad reference: force the fp32 kernels through a second process-independent route: env toggles are read
    # once, so compare against torch instead (fp64)
    print(sh, ' | '.join(res), flush=True)
    print('    |gk| %.3e |gb| %.3e' % (float(gk_tc.norm()), float(gb_tc.norm())), flush=True)
