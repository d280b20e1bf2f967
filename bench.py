#!/usr/bin/env python
"""Benchmark of the NLT UV-space hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # reference arithmetic on the host CPU cores

A "step" is one full training pass of the hot path over one synthetic batch:
Model.call (fused input gather + two-stream U-Net + UV->camera tail), L2 loss,
full backward to every weight gradient, (N>1: one NCCL all-reduce of the flat
gradient bucket) and the fused AMSGrad update.

Workload at every N: configs[1] of BASELINE.json -- dragon_specular network
(depth0 16, depth 256, k2 s2, LeakyReLU), 1024x1024 UV, batch 8 PER GPU (weak
scaling along the view x light axis).  Metric: UV texels/s forward+backward,
whole job.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'neural-light-transport_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = 'UV texels/sec fwd+bwd'
UNIT = 'texels/s'
# block-granularity algorithmic HBM bytes per texel, fwd+bwd
ALG_BYTES_PER_TEXEL = {5: 2210.0, 64: 2918.0}
ALG_FLOP_PER_TEXEL = {5: 3 * 13464.0, 64: 3 * 15352.0}


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d['hbm_gbs']), float(d.get('bf16_tflops_sustained', d.get('bf16_tflops', 1400.0))), 'measured'
    return 6650.0, 1400.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self.thread = None

    def start(self):
        def run():
            q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
                 'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
            while not self._stop.is_set():
                try:
                    out = subprocess.check_output(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                                   '--format=csv,noheader,nounits'], timeout=5).decode()
                    self.rows.append([x.strip() for x in out.strip().split(',')])
                except Exception:
                    pass
                self._stop.wait(0.2)
        self.thread = threading.Thread(target=run, daemon=True)
        self.thread.start()

    def stop(self):
        self._stop.set()
        if self.thread:
            self.thread.join(timeout=10)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i] == 'Active'})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def make_config(uv, cam):
    from util import io as ioutil
    return ioutil.make_config(uvh=uv, uvw=uv, imh=cam, imw=cam, loss='l2', depth0=16, depth=256, kernel=2, stride=2)


# ----------------------------------------------------------------------------
# reference arm: the restated reference arithmetic on the host CPU cores
# ----------------------------------------------------------------------------
def calibrate_cpu_threads():
    """Many-core hosts run this conv stack SLOWER with one torch thread per
    logical core (oversubscribed oneDNN primitives on small channel counts), so
    the reference arm picks the thread count that maximises its own throughput
    on a small probe and reports it as `cores`."""
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, cores) if c <= cores})
    fn, _ = cpu_reference_step_fn(256, 1)
    best, best_t = cands[0], None
    for c in cands:
        torch.set_num_threads(c)
        fn()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def cpu_reference_step_fn(uv, batch, c_extra=0):
    """Returns (fn, texels_per_call): one fwd+bwd of the oracle (torch CPU,
    fp32, all host threads) over `batch` samples at uv x uv."""
    from oracle import nlt_oracle as O
    from util import synth
    cfg = dict(depth0=16, depth=256, kernel=2, stride=2, norm='None', act='leakyrelu', pool='None', use_obs=True,
               skip_connect_base=True, imh=uv, imw=uv, uvh=uv, uvw=uv)
    params = O.init_params(cfg, c_query=5 + c_extra, seed=7, dtype=torch.float32)
    for v in params.values():
        v.requires_grad_(True)
    bt = synth.make_batch(batch, uv, uv, seed=1235, c_extra=c_extra)

    def fn():
        for v in params.values():
            v.grad = None
        loss = O.train_loss(params, cfg, bt, batch)
        loss.backward()
        return float(loss.detach())
    return fn, batch * uv * uv


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = calibrate_cpu_threads()
    fn, texels = cpu_reference_step_fn(args.uv, args.cpu_batch, args.c_extra)
    for _ in range(max(args.warmup, 1)):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = (time.perf_counter() - t0) / args.steps
    val = texels / dt
    sample = 'B=%d of the per-GPU batch %d at %dx%d UV, fwd+bwd, torch-CPU fp32 restatement of the TF2 path' % (
        args.cpu_batch, args.batch, args.uv, args.uv)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': max(args.warmup, 1), 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, 1),
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'host_logical_cores': os.cpu_count(),
                         'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    emit(line)


def workload_config(args, world):
    return {'workload': 'cfg2: dragon_specular net (depth0 16, depth 256, k2 s2, leakyrelu, K=1 obs), '
                        '%dx%d UV, %dx%d camera, batch %d per GPU, fwd+L2+bwd+AMSGrad' % (
                            args.uv, args.uv, args.uv, args.uv, args.batch),
            'query_channels': 5 + args.c_extra, 'global_batch': args.batch * world, 'uv': args.uv,
            'parallelism': 'dp%d' % world, 'cuda_graph': not getattr(args, 'no_graph', False),
            'l2_flush': 'inputs+activations per step (>2.5 GB) exceed the 126 MB L2; fresh batch buffers rotate'}


# ----------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------
def run_b200(args):
    import models
    import trainvali
    import nlt_native as nat
    import engine
    from util import synth

    strategy = trainvali.Strategy()
    world, rank = strategy.world, strategy.rank
    dev = torch.device('cuda', torch.cuda.current_device())
    cfg = make_config(args.uv, args.uv)
    Model = models.get_model_class('nlt')
    model = Model(cfg)
    model.register_trainable()
    opt = trainvali.Adam(learning_rate=1e-3, amsgrad=True)
    global_bs = args.batch * world
    # two rotating synthetic batches: device-resident copies and pinned host copies
    host = [synth.make_batch(args.batch, args.uv, args.uv, seed=1235 + 17 * rank + i, c_extra=args.c_extra, pin=True)
            for i in range(2)]
    resident = [tuple(t.to(dev) if torch.is_tensor(t) else t for t in b) for b in host]
    texels_step = args.batch * args.uv * args.uv * world

    graphed = None if args.no_graph else trainvali.GraphedTrainStep(strategy, model, opt, global_bs)

    def step(batch):
        if graphed is not None and not getattr(engine.PROF, 'force_eager', False):
            return graphed(batch)
        return trainvali.distributed_train_step(strategy, model, batch, opt, global_bs)

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize, device-timed, max over ranks."""
        strategy.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        strategy.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms) / steps

    # ---- device-resident arm ("value") ----
    for i in range(max(args.warmup, 3)):
        step(resident[i % 2])
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    l0 = nat.launch_count()
    tc0 = nat.tc_launch_count()
    ms_step = timed(lambda i: step(resident[i % 2]), args.steps)
    launches = nat.launch_count() - l0
    tc_launches = nat.tc_launch_count() - tc0
    if graphed is not None and graphed.graph is not None:
        # replayed kernels are not seen by the library's counter: add the captured ones per replay
        launches += args.steps * graphed.captured_launches
        tc_launches += args.steps * graphed.captured_tc_launches
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end arm: pinned host inputs -> H2D each step -> loss read back ----
    copy_stream = torch.cuda.Stream()
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0] if torch.is_tensor(t))
    state = {}

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            state['next'] = tuple(t.to(dev, non_blocking=True) if torch.is_tensor(t) else t for t in host[i % 2])
            state['ev'] = torch.cuda.Event()
            state['ev'].record(copy_stream)

    def e2e_step(i):
        torch.cuda.current_stream().wait_event(state['ev'])
        cur = state['next']
        prefetch(i + 1)            # overlaps the next step's H2D with this step's compute
        loss, _ = step(cur)
        prev = state.get('pending')
        state['pending'] = loss
        if prev is not None:
            state['loss'] = float(prev)   # D2H read of a step's result, one step behind the launch front

    prefetch(0)
    for i in range(2):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps)
    torch.cuda.current_stream().wait_event(state['ev'])

    # ---- roofline leg: per-call CUDA events (separate pass, not the timed value) ----
    roof = None
    # every rank runs these steps (they contain the gradient all-reduce); only rank 0 records events
    engine.PROF.enabled = (rank == 0)
    engine.PROF.force_eager = True
    engine.USE_SIDE_STREAM = False        # one stream: per-kernel times without concurrent-branch interference
    for i in range(2):
        step(resident[i % 2])
    engine.USE_SIDE_STREAM = True
    engine.PROF.force_eager = False
    if rank == 0:
        summ = engine.PROF.summary()
        engine.PROF.enabled = False
        total_ms = sum(v['ms'] for v in summ.values())
        top_label, top = max(summ.items(), key=lambda kv: kv[1]['ms'])
        peak_gbs, peak_tf, how = measured_peaks()
        ach = top['bytes'] / (top['ms'] * 1e-3) / 1e9
        cq = 5 + args.c_extra
        step_bytes = ALG_BYTES_PER_TEXEL.get(cq, 2210.0) * args.batch * args.uv * args.uv
        roof = {
            'bound': 'hbm', 'achieved': ach, 'peak': peak_gbs, 'unit': 'GB/s', 'frac': ach / peak_gbs,
            'traffic': ncu_traffic(top_label), 'peak_source': how + ' copy bandwidth (MEASURED_PEAKS.json hbm_gbs)'
            if how == 'measured' else 'fallback 6.65 TB/s (B200_PROFILING.md)',
            'kernel': top_label, 'kernel_launches_per_step': top['launches'] / 2,
            'kernel_ms_per_launch': top['ms'] / top['launches'],
            'kernel_alg_bytes_per_launch': top['bytes'] / top['launches'],
            'kernel_share_of_step': top['ms'] / total_ms,
            'step': {'alg_bytes': step_bytes, 'achieved_GBps': step_bytes / (ms_step * 1e-3) / 1e9,
                     'hbm_frac': step_bytes / (ms_step * 1e-3) / 1e9 / peak_gbs,
                     'alg_flops': ALG_FLOP_PER_TEXEL.get(cq, 40392.0) * args.batch * args.uv * args.uv,
                     'flop_frac_of_bf16_peak': ALG_FLOP_PER_TEXEL.get(cq, 40392.0) * args.batch * args.uv * args.uv
                     / (ms_step * 1e-3) / 1e12 / peak_tf},
            'top5': sorted(((k, round(v['ms'] / 2, 4)) for k, v in summ.items()), key=lambda kv: -kv[1])[:5],
        }
        if args.profile_out:
            rows = sorted(({'op': k, 'ms_per_step': v['ms'] / 2, 'launches_per_step': v['launches'] / 2,
                            'alg_MB_per_step': v['bytes'] / 2 / 1e6,
                            'GBps': v['bytes'] / (v['ms'] * 1e-3) / 1e9} for k, v in summ.items()),
                          key=lambda r: -r['ms_per_step'])
            json.dump({'ms_step': ms_step, 'sum_ms': total_ms / 2, 'rows': rows}, open(args.profile_out, 'w'), indent=1)

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = calibrate_cpu_threads()
        fn, texels = cpu_reference_step_fn(args.uv, args.cpu_batch, args.c_extra)
        fn()
        t0 = time.perf_counter()
        n = 2
        for _ in range(n):
            fn()
        dt = (time.perf_counter() - t0) / n
        cpu = {'value': texels / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': 'B=%d of the B=%d step at %dx%d UV, fwd+bwd, 1 warm-up + %d timed, torch-CPU fp32 '
                         'restatement of the TF2 path (TensorFlow is not installable offline)' % (
                             args.cpu_batch, args.batch, args.uv, args.uv, n)}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': texels_step / (ms_step * 1e-3), 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args, world),
            'e2e': {'value': texels_step / (ms_e2e * 1e-3), 'unit': UNIT, 'ms_per_step': ms_e2e,
                    'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4},
            'gpu_launches': launches, 'tcgen05_launches': tc_launches, 'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu,
        }
        emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


def ncu_traffic(label):
    """DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
    (profiles/ncu_traffic.json); None when that kernel has not been captured."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            ent = json.load(f).get(label)
        return float(ent['dram_bytes_per_launch']) if ent else None
    except (OSError, ValueError, KeyError):
        return None


_JSON_OUT = None


def _guard_stdout():
    """Keep stdout to the ONE JSON line of the contract: native libraries write banners to fd 1
    (NCCL prints its version there on the first communicator), so fd 1 is pointed at stderr and the
    result line goes to a private duplicate of the original stdout."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--uv', type=int, default=1024)
    ap.add_argument('--batch', type=int, default=8, help='per-GPU batch')
    ap.add_argument('--c-extra', dest='c_extra', type=int, default=0, help='59 -> cfg4 64-channel query stack')
    ap.add_argument('--cpu-batch', dest='cpu_batch', type=int, default=1)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', dest='no_graph', action='store_true', help='launch kernels eagerly (no CUDA graph)')
    ap.add_argument('--profile-out', dest='profile_out', default=None, help='write the per-op device-time table here')
    args = ap.parse_args()
    _guard_stdout()
    if args.impl == 'reference':
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit('bench.py: no CUDA device (the B200 path has no CPU fallback); '
                             'use --impl reference for the CPU arm')
        run_b200(args)


if __name__ == '__main__':
    main()
