"""Data-parallel training step -- host mirror of nlt/trainvali.py:254-325.

The reference replicates in-graph with tf.distribute.MirroredStrategy; here it
is one process per GPU (torchrun), the batch sharded along the view x light
axis, and ONE all-reduce(SUM) over the flat fp32 gradient bucket per step
(NCCL over NVLink/NVSwitch), followed by one fused AMSGrad launch.
"""
import os

import torch
import torch.distributed as dist

import nlt_native as nat


class Strategy:
    """Stand-in for tf.distribute.{OneDevice,Mirrored}Strategy
    (trainvali.py:254-264): rank / world size of the torchrun job."""

    def __init__(self, backend=None):
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        if self.world > 1 and not dist.is_initialized():
            if backend is None:
                backend = 'nccl' if torch.cuda.is_available() else 'gloo'
            if backend == 'nccl':
                torch.cuda.set_device(self.local_rank)
            dist.init_process_group(backend)
        if torch.cuda.is_available():
            torch.cuda.set_device(self.local_rank if self.world > 1 else torch.cuda.current_device())

    @property
    def num_replicas_in_sync(self):
        return self.world

    def all_reduce_sum_(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def barrier(self):
        if self.world > 1:
            dist.barrier()


def get_strategy(device='gpu'):
    if device == 'gpu':
        return Strategy()
    if device == 'cpu':
        # the reference offers OneDeviceStrategy('/cpu:0'); this hot path has
        # no CPU implementation by design
        raise NotImplementedError('device=cpu: the B200 hot path has no CPU fallback')
    raise NotImplementedError(device)


class Adam:
    """tf.keras.optimizers.Adam(learning_rate, amsgrad=True) over the model's
    flat parameter bucket (trainvali.py:122-127); Keras defaults b1 .9,
    b2 .999, eps 1e-7."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, amsgrad=False, clipnorm=None):
        if not amsgrad:
            raise NotImplementedError('only amsgrad=True is on the accelerated path')
        if clipnorm is not None:
            raise NotImplementedError('clipnorm (mgm > 0) is out of parity scope (SURVEY.md 8c)')
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.iterations = 0
        self.m = self.v = self.vhat = None

    def apply_gradients(self, grads_and_vars, grad_scale=1.0):
        gv = list(grads_and_vars)
        g0, v0 = gv[0]
        flat_g, flat_p = g0._base, v0._base
        if flat_g is None or flat_p is None or any(g._base is not flat_g or v._base is not flat_p for g, v in gv):
            raise ValueError('apply_gradients expects views of the model\'s flat parameter/gradient buckets')
        if self.m is None:
            self.m, self.v, self.vhat = (torch.zeros_like(flat_p) for _ in range(3))
        self.iterations += 1
        nat.check(nat.lib().nlt_amsgrad_step(
            nat.ptr(flat_p), nat.ptr(flat_g), nat.ptr(self.m), nat.ptr(self.v), nat.ptr(self.vhat),
            flat_p.numel(), self.iterations, self.lr, self.b1, self.b2, self.eps, grad_scale, nat.stream()))


def distributed_train_step(strategy, model, batch, optimizer, global_bs):
    """trainvali.py:267-290: per-replica train_step, gradient all-reduce,
    loss SUM.  `batch` is this rank's shard of the global batch."""
    assert model.trainable_registered, \
        "Register the trainable layers before using `trainable_variables`"

    def train_step(batch):
        pred, gt, loss_kwargs, partial_to_vis = model(batch, mode='train')
        loss_kwargs['keep_batch'] = True  # keep the batch dimension
        model.set_loss_grad_scale(1.0 / global_bs)
        per_example_loss = model.compute_loss(pred, gt, **loss_kwargs)
        weighted_loss = per_example_loss.sum() / global_bs   # tf.nn.compute_average_loss
        model.backward()
        strategy.all_reduce_sum_(model.flat_grads)           # ONE collective per step
        optimizer.apply_gradients(zip(model.gradients, model.trainable_variables))
        return weighted_loss, partial_to_vis

    weighted_loss, to_vis = train_step(batch)
    loss = strategy.all_reduce_sum_(weighted_loss.clone())
    return loss, to_vis


class GraphedTrainStep:
    """distributed_train_step with the forward + loss + backward of one replica
    captured ONCE in a CUDA graph (shapes are static; every activation, gradient
    and scratch buffer lives in the graph's private pool, so TMA descriptors and
    kernel arguments stay valid across replays).  Per step: copy the batch into
    the static input buffers, replay, all-reduce the flat gradient bucket, one
    fused AMSGrad launch (its bias-corrected lr is a host scalar, so it stays
    outside the graph).  Same semantics as nlt/trainvali.py:267-290."""

    def __init__(self, strategy, model, optimizer, global_bs):
        self.strategy, self.model, self.optimizer, self.global_bs = strategy, model, optimizer, global_bs
        self.graph = None
        self.static_batch = None
        self.loss = None
        self.to_vis = None

    def _body(self, batch):
        model = self.model
        pred, gt, loss_kwargs, to_vis = model(batch, mode='train')
        loss_kwargs['keep_batch'] = True
        model.set_loss_grad_scale(1.0 / self.global_bs)
        per_example_loss = model.compute_loss(pred, gt, **loss_kwargs)
        weighted_loss = per_example_loss.sum() / self.global_bs
        model.backward()
        return weighted_loss, to_vis

    def _capture(self, batch):
        dev = self.model.device
        self.static_batch = tuple(t.to(dev, torch.float32).contiguous().clone() if torch.is_tensor(t) else t
                                  for t in batch)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):      # warm-up: lazy build, one-time CUDA attribute calls, scratch sizing
            for _ in range(2):
                self._body(self.static_batch)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        l0, t0 = nat.launch_count(), nat.tc_launch_count()
        with torch.cuda.graph(self.graph):
            self.loss, self.to_vis = self._body(self.static_batch)
        # kernels of this library inside ONE replay (the counters only see the capture)
        self.captured_launches = nat.launch_count() - l0
        self.captured_tc_launches = nat.tc_launch_count() - t0

    def __call__(self, batch):
        assert self.model.trainable_registered, \
            "Register the trainable layers before using `trainable_variables`"
        if self.graph is None:
            self._capture(batch)
        if batch is not self.static_batch:
            for dst, src in zip(self.static_batch, batch):
                if torch.is_tensor(dst):
                    dst.copy_(src, non_blocking=True)
        self.graph.replay()
        self.strategy.all_reduce_sum_(self.model.flat_grads)
        self.optimizer.apply_gradients(zip(self.model.gradients, self.model.trainable_variables))
        loss = self.strategy.all_reduce_sum_(self.loss.clone())
        return loss, self.to_vis


def distributed_vali_step(strategy, model, batch, global_bs):
    """trainvali.py:296-312."""
    pred, gt, loss_kwargs, to_vis = model(batch, mode='vali')
    loss_kwargs['keep_batch'] = True
    per_example_loss = model.compute_loss(pred, gt, **loss_kwargs)
    weighted_loss = per_example_loss.sum() / global_bs
    loss = strategy.all_reduce_sum_(weighted_loss.clone())
    return loss, to_vis


def batch_source(config, strategy, steps, device='cuda', mode='train', seed=None):
    """This rank's shards of `steps` global batches (trainvali.py:72-90).

    Real data when `<data_root>.json` exists: `datasets.get_dataset_class(config dataset)` -> pipeline ->
    `.shard(world, rank)` (the reference's `experimental_distribute_dataset`), cycling over epochs; the tensors
    arrive pinned and are copied to `device` asynchronously.  Otherwise (no NLT data on this machine) synthetic
    batches of the configured shape (util/synth.py)."""
    import os
    from util import synth
    bs = config.getint('DEFAULT', 'bs')
    data_root = config.get('DEFAULT', 'data_root', fallback='')
    if data_root and os.path.exists(data_root.rstrip('/') + '.json'):
        import datasets
        geti = lambda k, d: config.getint('DEFAULT', k, fallback=d)
        pre, par = geti('prefetch_buffer_size', -1), geti('n_map_parallel_calls', -1)
        ds = datasets.get_dataset_class(config.get('DEFAULT', 'dataset', fallback='nlt'))(
            config, mode, shuffle_buffer_size=geti('shuffle_buffer_size', 64),
            prefetch_buffer_size=None if pre < 0 else pre, n_map_parallel_calls=None if par < 0 else par)
        pipe = ds.build_pipeline(seed=seed, no_batch=config.getboolean('DEFAULT', 'no_batch', fallback=False))
        pipe = pipe.shard(strategy.world, strategy.rank)
        done = 0
        while done < steps:
            n_before = done
            for batch in pipe:
                yield tuple(t.to(device, non_blocking=True) if torch.is_tensor(t) else t for t in batch)
                done += 1
                if done >= steps:
                    return
            if done == n_before:
                raise RuntimeError('dataset pipeline produced no batch')
        return
    local_bs = max(bs // strategy.world, 1)
    for step in range(steps):
        yield synth.make_batch(local_bs, config.getint('DEFAULT', 'uvh'), config.getint('DEFAULT', 'imh'),
                               seed=1234 + step * strategy.world + strategy.rank, device=device)


def main(argv=None):
    """Same CLI flags as the reference (--config --debug --device).  Trains from the on-disk dataset named by
    the config when it exists (datasets/nlt.py), else on synthetic batches of the configured shape."""
    import argparse
    import models
    from util import io as ioutil
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='dragon_specular.ini')
    ap.add_argument('--debug', action='store_true')
    ap.add_argument('--device', default='gpu')
    ap.add_argument('--steps', type=int, default=4)
    args = ap.parse_args(argv)
    strategy = get_strategy(args.device)
    config = ioutil.read_config(args.config)
    Model = models.get_model_class(config.get('DEFAULT', 'model'))
    model = Model(config)
    model.register_trainable()
    optimizer = Adam(learning_rate=config.getfloat('DEFAULT', 'lr'), amsgrad=True)
    global_bs = config.getint('DEFAULT', 'bs')
    steps = 1 if args.debug else args.steps
    for step, batch in enumerate(batch_source(config, strategy, steps, device='cuda')):
        loss, _ = distributed_train_step(strategy, model, batch, optimizer, global_bs)
        if strategy.rank == 0:
            print('step %d loss %.6f' % (step, float(loss)))


if __name__ == '__main__':
    main()
