"""Checkpoints of the hot path's trainable state -- the role of `tf.train.Checkpoint(step=, optimizer=, net=model)` +
`tf.train.CheckpointManager` in nlt/trainvali.py:134-141 and of `restore_model` in nlt/nlt_test.py:61-75.

One `.npz` per checkpoint, written atomically.  Keys follow the reference's object-graph names
(`register_trainable()` publishes every layer as `net_<network>_layer<i>`, nlt/models/base.py:79-101):

    net/net_query_layer3/conv0/kernel      Keras layout (kh,kw,Ci,Co) / (kh,kw,Co,Ci), float32
    net/net_query_layer3/conv0/bias
    optimizer/m/<same suffix> , optimizer/v/... , optimizer/vhat/...      AMSGrad slots
    optimizer/iterations , step

so that the file is independent of how the flat parameter bucket is laid out.  The TF tensor-bundle reader for the
released NLT checkpoints lives in util/tf_ckpt.py and produces the same key set.
"""
import os
import re

import numpy as np
import torch


def _named(model):
    """[(key stem, ConvLayer)] in registration order: 'net/net_query_layer3/conv0'."""
    assert model.trainable_registered, \
        "Register the trainable layers to have them tracked by the checkpoint"
    out = []
    for net_name, net in model.net.items():
        for li, blk in enumerate(net.layers):
            for suffix, c in blk.param_layers():          # '0', '1', and '0.norm' / '1.norm' for norm = instance
                ci, _, nm = suffix.partition('.')
                out.append(('net/net_%s_layer%d/%s%s' % (net_name, li, 'norm' if nm else 'conv', ci), c))
    return out


def _slot_views(model, flat):
    """Per-layer views of a flat optimiser slot that shares the parameter bucket's layout."""
    views = {}
    bucket = model.bucket
    for (stem, c), (ko, ks, bo, bs) in zip(_named_in_bucket_order(model), bucket.slices):
        views[stem + '/kernel'] = flat[ko:ko + ks].view(c.kernel.shape)
        views[stem + '/bias'] = flat[bo:bo + bs].view(c.bias.shape)
    return views


def _named_in_bucket_order(model):
    by_id = {id(c): stem for stem, c in _named(model)}
    return [(by_id[id(c)], c) for c in model.bucket.layers]


def state_dict(model, optimizer=None, step=0):
    if model.bucket is None:
        raise RuntimeError('the model has no parameters yet (call model.build() or run one batch first)')
    out = {'step': np.int64(step)}
    for stem, c in _named(model):
        out[stem + '/kernel'] = c.kernel.detach().cpu().numpy()
        out[stem + '/bias'] = c.bias.detach().cpu().numpy()
    if optimizer is not None and optimizer.m is not None:
        out['optimizer/iterations'] = np.int64(optimizer.iterations)
        for slot in ('m', 'v', 'vhat'):
            for k, v in _slot_views(model, getattr(optimizer, slot)).items():
                out['optimizer/%s/%s' % (slot, k[len('net/'):])] = v.detach().cpu().numpy()
    return out


def save(path, model, optimizer=None, step=0):
    """Writes `<path>` (a '.npz' suffix is appended when missing) atomically; returns the final path."""
    if not path.endswith('.npz'):
        path += '.npz'
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    tmp = path + '.tmp.npz'
    np.savez(tmp, **state_dict(model, optimizer, step))
    os.replace(tmp, path)
    return path


def load_state(model, state, optimizer=None, expect_partial=False):
    """Copies a {key: array} state (ours or util/tf_ckpt.py's) into the model (and optimiser slots when present).
    The model must be built (channel counts are fixed by its first batch / model.build()).  Returns `step`."""
    if model.bucket is None:
        raise RuntimeError('build the model (model.build(c_query, c_obs)) before restoring a checkpoint')
    missing = []
    for stem, c in _named(model):
        for leaf, dst in (('kernel', c.kernel), ('bias', c.bias)):
            key = stem + '/' + leaf
            if key not in state:
                missing.append(key)
                continue
            src = torch.as_tensor(np.asarray(state[key]), dtype=torch.float32)
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError('%s: checkpoint shape %s, model shape %s' % (key, tuple(src.shape), tuple(dst.shape)))
            dst.copy_(src.to(dst.device))
    if missing and not expect_partial:
        raise KeyError('checkpoint lacks %d model tensors, e.g. %s' % (len(missing), missing[0]))
    if optimizer is not None and 'optimizer/iterations' in state:
        flat = model.flat_params
        slots = {}
        for slot in ('m', 'v', 'vhat'):
            buf = torch.zeros_like(flat)
            for k, dst in _slot_views(model, buf).items():
                key = 'optimizer/%s/%s' % (slot, k[len('net/'):])
                if key in state:
                    dst.copy_(torch.as_tensor(np.asarray(state[key]), dtype=torch.float32).to(dst.device))
                elif not expect_partial:
                    raise KeyError(key)
            slots[slot] = buf
        optimizer.load_state(int(state['optimizer/iterations']), slots['m'], slots['v'], slots['vhat'])
    return int(state['step']) if 'step' in state else 0


def restore(path, model, optimizer=None, expect_partial=False):
    """`ckpt.restore(path)` of the reference: our .npz, or a TF tensor-bundle prefix (`.../ckpt-43`)."""
    if os.path.exists(path) and path.endswith('.npz'):
        with np.load(path) as z:
            state = {k: z[k] for k in z.files}
    elif os.path.exists(path + '.npz'):
        with np.load(path + '.npz') as z:
            state = {k: z[k] for k in z.files}
    elif os.path.exists(path + '.index'):
        from util import tf_ckpt
        state = tf_ckpt.load_nlt_state(path)
        expect_partial = True          # the reference restores with .expect_partial() (nlt_test.py:73)
    else:
        raise FileNotFoundError(path)
    return load_state(model, state, optimizer, expect_partial)


class CheckpointManager:
    """tf.train.CheckpointManager(ckpt, ckptdir, max_to_keep): numbered `ckpt-<n>.npz` files, newest kept."""

    def __init__(self, ckptdir, max_to_keep=None):
        self.dir = ckptdir
        self.max_to_keep = max_to_keep

    def _numbered(self):
        if not os.path.isdir(self.dir):
            return []
        found = []
        for f in os.listdir(self.dir):
            m = re.fullmatch(r'ckpt-(\d+)\.npz', f)
            if m:
                found.append((int(m.group(1)), os.path.join(self.dir, f)))
        return sorted(found)

    @property
    def latest_checkpoint(self):
        found = self._numbered()
        return found[-1][1] if found else None

    def save(self, model, optimizer=None, step=0):
        found = self._numbered()
        n = found[-1][0] + 1 if found else 1
        path = save(os.path.join(self.dir, 'ckpt-%d' % n), model, optimizer, step)
        if self.max_to_keep:
            for _, old in self._numbered()[:-self.max_to_keep]:
                os.remove(old)
        return path

    def restore_latest(self, model, optimizer=None):
        """util/io.py:32-37 `restore`: latest checkpoint when there is one, else start from scratch (returns None)."""
        latest = self.latest_checkpoint
        if latest is None:
            return None
        return restore(latest, model, optimizer)
