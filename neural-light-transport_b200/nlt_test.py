"""Test-time path -- host mirror of nlt/nlt_test.py:78-127: the observation
encoder is run once over training samples, its per-level mean becomes a
K-independent feature cache that overrides the observation stream at inference."""
import torch

import nlt_native as nat
from engine import Act, Seg


def restore_model(config, ckpt, c_query=5, c_obs=3):
    """nlt/nlt_test.py:61-75: Model(config) -> register_trainable() -> restore the checkpoint (`.expect_partial()`).
    `ckpt`: a checkpoint written by util/ckpt.py (`.../ckpt-43.npz`) or the prefix of a TensorFlow checkpoint of the
    reference (`.../ckpt-43`, read by util/tf_ckpt.py).  c_query / c_obs: input channel counts of the two streams
    (Keras infers them from the first batch; a checkpoint restore needs them up front)."""
    import models
    from util import ckpt as ckptutil
    Model = models.get_model_class(config.get('DEFAULT', 'model'))
    model = Model(config)
    model.register_trainable()
    assert model.trainable_registered, (
        "Register the trainable layers to have them restored from the "
        "checkpoint")
    model.build(c_query, c_obs)
    ckptutil.restore(ckpt, model, expect_partial=True)
    return model


def extract_feat(model, datapipe, n_obs_batches=1):
    """nlt_test.py:97-127.  datapipe yields the 11-tuple batches; `n_obs_batches` mirrors the reference flag
    (default 1; <= 0: all batches) and is applied BEFORE the pipeline is iterated (`datapipe.take(n)`)."""
    obs_feat_extractor = model.net['obs']
    if not obs_feat_extractor.layers or not all(blk.built for blk in obs_feat_extractor.layers):
        raise RuntimeError('extract_feat needs a built model: restore a checkpoint (restore_model) or run / build '
                           'the model first -- an unbuilt model would silently yield features of random weights')
    if n_obs_batches > 0:
        if hasattr(datapipe, 'take'):
            datapipe = datapipe.take(n_obs_batches)
        else:
            import itertools
            datapipe = itertools.islice(iter(datapipe), n_obs_batches)
    lib = nat.lib()
    sums, count = None, 0
    for batch in datapipe:
        _, base, _, _, _, rgb, _, _, _, _, _ = batch
        dev = model.device
        from models.nlt import _dev_tensor
        rgb, base = _dev_tensor(rgb, dev), _dev_tensor(base, dev)
        # Forward through the observation path; x = rgb - base is formed inside the first conv's operand loader
        segs = [Seg(Act(rgb), sub=base)]
        feat = []
        for layer in obs_feat_extractor.layers:
            y = layer.forward_segs(segs)
            feat.append(y.t)
            segs = [Seg(y)]
        # running per-level sum over all samples instead of the reference's concat-then-mean (same value up to
        # fp32 summation order, without holding every sample): nlt_kmean_fwd with K = batch samples
        n = feat[0].shape[0]
        if sums is None:
            sums = [torch.zeros((1,) + tuple(f.shape[1:]), dtype=torch.float32, device=dev) for f in feat]
        for s, f in zip(sums, feat):
            per = f.shape[1] * f.shape[2] * f.shape[3]
            nat.check(lib.nlt_ksum_acc(nat.ptr(f), n, per, nat.ptr(s), nat.stream()))
        count += n
    if sums is None:
        raise RuntimeError('the observation pipeline produced no batch')
    # Each element is 1xHxWxC
    for s in sums:
        nat.check(lib.nlt_scale(nat.ptr(s), s.numel(), 1.0 / count, nat.stream()))
    return sums


def infer(model, datapipe, feat_agg, on_batch=None):
    """nlt_test.py:78-94: feat_agg overrides the observation stream.  The
    reference tiles it to the batch size; here the [1,h,w,c] maps are
    broadcast inside the kernels (never tiled in memory)."""
    outs = []
    for batch_i, batch in enumerate(datapipe):
        pred_camspc, _, _, to_vis = model.call(batch, 'test', obs_override=feat_agg)
        if on_batch is not None:
            on_batch(batch_i, to_vis)
        outs.append(pred_camspc)
    return outs


def get_config_ini(ckpt):
    """nlt_test.py:47-48: <outroot>/<xname>/checkpoints/ckpt-N -> <outroot>/<xname>.ini"""
    return '/'.join(ckpt.split('/')[:-2]) + '.ini'


def make_datapipe(mode, config):
    """nlt_test.py:51-58."""
    import datasets
    Dataset = datasets.get_dataset_class(config.get('DEFAULT', 'dataset'))
    dataset = Dataset(config, mode)
    return dataset.build_pipeline(no_batch=config.getboolean('DEFAULT', 'no_batch'))


def main(argv=None):
    """Same flags as the reference (nlt_test.py:33-42): --ckpt --batch_size_override --n_obs_batches --fps."""
    import argparse
    from glob import glob
    from os.path import basename, join
    from util import io as ioutil
    ap = argparse.ArgumentParser()
    ap.add_argument('--ckpt', required=True, help="path to checkpoint (prefix only, e.g., '/path/to/ckpt-43')")
    ap.add_argument('--batch_size_override', type=int, default=None)
    ap.add_argument('--n_obs_batches', type=int, default=1,
                    help='number of observation batches used for the observation path')
    ap.add_argument('--fps', type=int, default=24)
    args = ap.parse_args(argv)
    config_ini = get_config_ini(args.ckpt)
    config = ioutil.read_config(config_ini)
    if args.batch_size_override is not None:
        config.set('DEFAULT', 'bs', str(args.batch_size_override))
    model = restore_model(config, args.ckpt)
    datapipe_train = make_datapipe('train', config)
    datapipe_test = make_datapipe('test', config)
    feat_agg = extract_feat(model, datapipe_train, args.n_obs_batches)
    outroot = join(config_ini[:-4], 'vis_test', basename(args.ckpt) + '_pred')

    def on_batch(batch_i, to_vis):
        model.vis_batch(to_vis, join(outroot, 'batch{i:09d}'.format(i=batch_i)), 'test')
    infer(model, datapipe_test, feat_agg, on_batch)
    batch_vis_dirs = sorted(glob(join(outroot, '*')))
    view_at = model.compile_batch_vis(batch_vis_dirs, outroot.rstrip('/'), 'test', fps=args.fps)
    print('Compilation available for viewing at\n\t%s' % view_at)


if __name__ == '__main__':
    main()
