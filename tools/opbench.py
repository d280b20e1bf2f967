#!/usr/bin/env python
"""Per-layer micro-benchmark: device time of ONE conv layer's forward / input-gradient / weight-gradient launches
at the benchmark shapes (dragon_specular network, 1024x1024 UV, batch 8), through the same engine + C ABI the model
uses.  Meant for A/B runs of kernel variants inside a single GPU session, e.g.

    python tools/opbench.py --layers query.1.1 obs.2.1 query.12.0 --opt dconv_wide32=1
    NLT_DISABLE_TC=1 python tools/opbench.py --layers query.3.0

Times are CUDA-event averages over --iters launches after --warmup (inputs far larger than L2, so no flush is
needed for the full-resolution layers; deep layers fit in L2 and are reported as such).  `--list` prints the layer
table without touching the GPU.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'neural-light-transport_b200'))
sys.path.insert(0, ROOT)


def layer_table(uv=1024, batch=8, depth0=16, depth=256, k=2, s=2, c_query=5, c_obs=3):
    """[(name, kind, k, stride, H_in, [segment channels], C_out)] of every conv of the two streams, replaying the
    wiring of models.nlt (level-0 1x1 convs, down blocks with the obs aggregate concatenated, up blocks with skips)."""
    from util.net import gen_feat_n
    n_feat = gen_feat_n(depth0, depth)
    rows = []
    h = uv
    rows.append(('obs.0.0', 'conv', 1, 1, h, [c_obs], n_feat[0]))
    rows.append(('query.0.0', 'conv', 1, 1, h, list(c_query) if isinstance(c_query, (list, tuple)) else [c_query], n_feat[0]))
    x = [n_feat[0], n_feat[0]]              # query stream input of the next layer: query_y (+) observation aggregate
    obs_c = n_feat[0]
    skips = [list(x)]                       # pushed after every contracting layer (models/nlt.py:171-177)
    prev, level = 0, 1
    for n in n_feat[:-1]:
        if n >= prev:                       # contracting block
            rows.append(('obs.%d.0' % level, 'conv', k, s, h, [obs_c], n))
            rows.append(('query.%d.0' % level, 'conv', k, s, h, list(x), n))
            h //= s
            rows.append(('obs.%d.1' % level, 'conv', k, 1, h, [n], n))
            rows.append(('query.%d.1' % level, 'conv', k, 1, h, [n], n))
            obs_c = n
            x = [n, n]
            skips.append(list(x))
        else:                               # expanding block: current input (+) popped skip (models/nlt.py:182-198)
            rows.append(('query.%d.0' % level, 'deconv', k, s, h, list(x) + skips.pop(), n))
            h *= s
            rows.append(('query.%d.1' % level, 'deconv', k, 1, h, [n], n))
            x = [n]
        prev = n
        level += 1
    rows.append(('query.%d.0' % level, 'conv', 1, 1, h, list(x) + skips.pop(), n_feat[-1]))
    return [(name, kind, kk, ss, hh, segc, cout, batch) for name, kind, kk, ss, hh, segc, cout in rows]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--layers', nargs='*', default=None, help='layer names (default: all); see --list')
    ap.add_argument('--list', action='store_true')
    ap.add_argument('--uv', type=int, default=1024)
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--opt', nargs='*', default=[], help='nlt_set_option pairs, e.g. dconv_wide32=1 tc=0')
    ap.add_argument('--graph', action='store_true',
                    help='GPU-only times: capture the launches in a CUDA graph and time replays (eager CUDA-event times of '
                         'microsecond kernels include the host latency between launches)')
    ap.add_argument('--cq-segs', dest='cq_segs', type=int, nargs='*', default=[3, 1, 1],
                    help='channel counts of the query input sources (cfg4: 3 60 1)')
    args = ap.parse_args()
    table = layer_table(args.uv, args.batch, c_query=args.cq_segs)
    if args.list:
        for r in table:
            print('%-12s %-6s k%d s%d  H_in %4d  segments %-14s -> %d' % (r[0], r[1], r[2], r[3], r[4], r[5], r[6]))
        return
    import torch
    import engine
    import nlt_native as nat
    for pair in args.opt:
        name, val = pair.split('=')
        nat.set_option(name, int(val))
    dev = torch.device('cuda')
    engine.USE_SIDE_STREAM = False
    want = set(args.layers) if args.layers else None
    if args.graph:
        print('%-12s %10s %10s   (ms per layer, CUDA-graph replay: fwd | fwd + dgrad + wgrad)' % ('layer', 'fwd', 'fwd+bwd'))
    else:
        print('%-12s %10s %10s %10s   (ms per launch group; dgrad = all segments)' % ('layer', 'fwd', 'dgrad', 'wgrad'))
    for name, kind, k, s, h, segc, cout, batch in table:
        if want is not None and name not in want:
            continue
        L = engine.ConvLayer(kind, k, s, cout, None if name.endswith('13.0') or name.endswith('.0.0') else 'leakyrelu')
        L.build(sum(segc), dev, torch.Generator().manual_seed(1))
        xs = [torch.randn(batch, h, h, c, device=dev) for c in segc]
        if args.graph:
            reps = 4

            def body(bwd):
                for _ in range(reps):
                    acts = [engine.Act(x, act='leakyrelu', needs_grad=True) for x in xs]
                    tape = engine.Tape() if bwd else None
                    y = L.forward([engine.Seg(a) for a in acts], tape)
                    if bwd:
                        y.grad = torch.ones_like(y.t)
                        tape.backward()
            res = []
            for bwd in (False, True):
                ws = (engine.Workspace(), engine.Workspace())
                with engine.use_workspaces(*ws):
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        body(bwd)
                    torch.cuda.current_stream().wait_stream(side)
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        body(bwd)
                for _ in range(args.warmup):
                    g.replay()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.iters):
                    g.replay()
                e1.record()
                torch.cuda.synchronize()
                res.append(e0.elapsed_time(e1) / args.iters / reps)
                del g
            print('%-12s %10.4f %10.4f' % (name, res[0], res[1]))
            continue
        times = {'fwd': 0.0, 'dgrad': 0.0, 'wgrad': 0.0}
        for it in range(args.warmup + args.iters):
            engine.PROF.records = []
            engine.PROF.enabled = True
            acts = [engine.Act(x, act='leakyrelu', needs_grad=True) for x in xs]
            tape = engine.Tape()
            y = L.forward([engine.Seg(a) for a in acts], tape)
            y.grad = torch.ones_like(y.t)
            tape.backward()
            torch.cuda.synchronize()
            summ = engine.PROF.summary()
            engine.PROF.enabled = False
            if it >= args.warmup:
                for key, v in summ.items():
                    times[key.split(' ')[0]] += v['ms']
            del acts, y
        n = max(1, args.iters)
        print('%-12s %10.3f %10.3f %10.3f' % (name, times['fwd'] / n, times['dgrad'] / n, times['wgrad'] / n))


if __name__ == '__main__':
    main()
