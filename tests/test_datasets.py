"""Input pipeline: on-disk NLT scene -> the model's 11-tuple.  CPU only.

A tiny scene is written to a temp dir in the reference's layout (status JSON next to the data root, PNG maps,
float16 uv2cam, nn.json) and read back through `datasets.get_dataset_class('nlt')`.  Expected values are
spelled out by hand where the arithmetic is exact (k/255, k/65535, identity resize, float16 warp) and through
cv2's own bilinear resize where it is not."""
import configparser
import json
import os

import cv2
import numpy as np
import pytest
import torch
from PIL import Image

CAMS, LIGHTS = ('cam0', 'cam1'), ('light0', 'light1')
UV_NATIVE, UVH = 16, 8            # stored 16x16 UV maps, consumed at 8x8
CAM_NATIVE, IMH, IMW = (12, 10), 6, 5


def _rng_img(rng, h, w, c=None, dtype=np.uint8):
    hi = np.iinfo(dtype).max
    shape = (h, w) if c is None else (h, w, c)
    return rng.integers(0, hi + 1, size=shape, dtype=dtype)


def _write_example(root, id_, rng, nn=('cam0', 'light0'), with_gt=True):
    d = os.path.join(root, id_)
    os.makedirs(d)
    arrs = {}
    arrs['diffuse'] = _rng_img(rng, UV_NATIVE, UV_NATIVE, 4)               # RGBA: alpha must be dropped
    arrs['cvis'] = _rng_img(rng, UV_NATIVE, UV_NATIVE)
    arrs['lvis'] = _rng_img(rng, UV_NATIVE, UV_NATIVE)
    Image.fromarray(arrs['diffuse']).save(os.path.join(d, 'diffuse.png'))
    Image.fromarray(arrs['cvis']).save(os.path.join(d, 'cvis.png'))
    Image.fromarray(arrs['lvis']).save(os.path.join(d, 'lvis.png'))
    arrs['uv2cam'] = rng.random((IMH, IMW, 2)).astype(np.float16)
    np.save(os.path.join(d, 'uv2cam.npy'), arrs['uv2cam'])
    rel = {k: os.path.join(id_, k + ('.npy' if k == 'uv2cam' else '.png')) for k in ('diffuse', 'cvis', 'lvis', 'uv2cam')}
    if with_gt:
        arrs['rgb'] = _rng_img(rng, UV_NATIVE, UV_NATIVE, 3)
        Image.fromarray(arrs['rgb']).save(os.path.join(d, 'rgb.png'))
        arrs['rgb_camspc'] = _rng_img(rng, CAM_NATIVE[0], CAM_NATIVE[1], 4)
        Image.fromarray(arrs['rgb_camspc']).save(os.path.join(d, 'rgb_camspc.png'))
        rel['rgb'] = os.path.join(id_, 'rgb.png')
        rel['rgb_camspc'] = os.path.join(id_, 'rgb_camspc.png')
    with open(os.path.join(d, 'nn.json'), 'w') as h:
        json.dump({'cam': nn[0], 'light': nn[1]}, h)
    rel['nn'] = os.path.join(id_, 'nn.json')
    rel['complete'] = True
    return rel, arrs


@pytest.fixture()
def scene(tmp_path):
    rng = np.random.default_rng(7)
    root = str(tmp_path / 'scene')
    os.makedirs(root)
    status, arrs = {}, {}
    i = 0
    for cam in CAMS:
        for light in LIGHTS:
            id_ = 'trainvali_%09d_%s_%s' % (i, cam, light)
            status[id_], arrs[id_] = _write_example(root, id_, rng)
            i += 1
    # an incomplete example (skipped), and a test view whose neighbour does not exist
    status['trainvali_%09d_cam0_light9' % i] = dict(status['trainvali_000000000_cam0_light0'], complete=False)
    tid = 'test_000000000_camT_lightT'
    status[tid], arrs[tid] = _write_example(root, tid, rng, nn=('camX', 'lightX'), with_gt=False)
    with open(root + '.json', 'w') as h:
        json.dump(status, h)
    cfg = configparser.ConfigParser()
    cfg['DEFAULT'] = dict(data_root=root, holdout_cam='cam1', holdout_light='light1', imh=str(IMH), imw=str(IMW),
                          uvh=str(UVH), uvw=str(UVH), bs='2', cache='False')
    return cfg, arrs


def _dataset(cfg, mode, **kw):
    import datasets
    return datasets.get_dataset_class('nlt')(cfg, mode, **kw)


def test_mode_split_holdout_and_incomplete(scene):
    cfg, _ = scene
    train, vali, test = (_dataset(cfg, m) for m in ('train', 'vali', 'test'))
    assert sorted(vali.files) == ['trainvali_000000003_cam1_light1']
    assert sorted(train.files) == ['trainvali_00000000%d_cam%d_light%d' % (i, i // 2, i % 2) for i in range(3)]
    assert test.files == ['test_000000000_camT_lightT']
    assert all('light9' not in f for f in train.files + vali.files)       # incomplete example skipped
    with pytest.raises(ValueError):
        _dataset(cfg, 'predict')
    cfg2 = configparser.ConfigParser()
    cfg2['DEFAULT'] = dict(cfg['DEFAULT'], data_root=cfg['DEFAULT']['data_root'] + '_missing')
    with pytest.raises(FileNotFoundError):
        _dataset(cfg2, 'train')


def test_example_arithmetic(scene):
    cfg, arrs = scene
    ds = _dataset(cfg, 'train')
    id_ = 'trainvali_000000001_cam0_light1'
    ex = ds._process_example_precache(id_)
    eid, base, cvis, lvis, warp, rgb, rgb_camspc, nn_id, nn_base, nn_rgb, nn_rgb_camspc = ex
    a = arrs[id_]
    assert eid == id_.encode() and nn_id == b'trainvali_000000000_cam0_light0'
    assert all(t.dtype == np.float32 for t in ex if isinstance(t, np.ndarray))
    assert base.shape == (UVH, UVH, 3) and cvis.shape == (UVH, UVH, 1) and lvis.shape == (UVH, UVH, 1)
    assert rgb.shape == (UVH, UVH, 3) and rgb_camspc.shape == (IMH, IMW, 3) and nn_rgb_camspc.shape == (IMH, IMW, 3)
    # normalise by the dtype maximum in float64, drop alpha, cv2 bilinear resize, then float32
    want = cv2.resize(a['diffuse'][:, :, :3].astype(float) / 255, (UVH, UVH)).astype(np.float32)
    np.testing.assert_array_equal(base, want)
    want = cv2.resize(a['rgb_camspc'][:, :, :3].astype(float) / 255, (IMW, IMH)).astype(np.float32)
    np.testing.assert_array_equal(rgb_camspc, want)
    np.testing.assert_array_equal(cvis[:, :, 0], cv2.resize(a['cvis'].astype(float) / 255, (UVH, UVH)).astype(np.float32))
    # the warp is taken as stored (float16 values, its own resolution): never resized
    assert warp.shape == (IMH, IMW, 2)
    np.testing.assert_array_equal(warp, a['uv2cam'].astype(np.float32))
    # the neighbour's maps are the neighbour example's own maps
    nb = ds._process_example_precache(nn_id.decode())
    np.testing.assert_array_equal(nn_base, nb[1])
    np.testing.assert_array_equal(nn_rgb, nb[5])
    np.testing.assert_array_equal(nn_rgb_camspc, nb[6])


def test_identity_resize_gives_exact_k_over_255(scene):
    cfg, arrs = scene
    cfg['DEFAULT']['uvh'] = str(UV_NATIVE)
    ds = _dataset(cfg, 'train')
    id_ = 'trainvali_000000000_cam0_light0'
    ex = ds._process_example_precache(id_)
    np.testing.assert_array_equal(ex[1], (arrs[id_]['diffuse'][:, :, :3].astype(np.float64) / 255).astype(np.float32))
    np.testing.assert_array_equal(ex[3][:, :, 0], (arrs[id_]['lvis'].astype(np.float64) / 255).astype(np.float32))


def test_uint16_maps_normalise_by_65535(scene, tmp_path):
    from datasets import nlt as dsnlt
    p = str(tmp_path / 'vis16.png')
    a = np.array([[0, 65535], [32768, 1]], dtype=np.uint16)
    Image.fromarray(a).save(p)
    got = dsnlt._unit_range(dsnlt._read_uint_image(p))
    np.testing.assert_array_equal(got, a.astype(np.float64) / 65535)
    with pytest.raises(TypeError):
        dsnlt._unit_range(np.zeros((2, 2), dtype=np.int32))
    with pytest.raises(ValueError):
        dsnlt._fit(np.zeros((4, 4)))


def test_test_mode_placeholders_and_missing_neighbour(scene):
    cfg, _ = scene
    ds = _dataset(cfg, 'test')
    ex = ds._process_example_precache('test_000000000_camT_lightT')
    assert ex[7] == b'incomplete-data_camX_lightX'
    assert ex[5].shape == (UVH, UVH, 3) and not ex[5].any()              # rgb placeholder
    assert ex[6].shape == (IMH, IMW, 3) and not ex[6].any()              # rgb_camspc placeholder
    assert not ex[8].any() and not ex[9].any() and not ex[10].any()      # black neighbour
    assert ex[1].any()                                                   # the diffuse base is real data


def test_ambiguous_neighbour_raises(scene):
    cfg, _ = scene
    ds = _dataset(cfg, 'train')
    ds.data_paths['trainvali_000000099_cam0_light0'] = ds.data_paths['trainvali_000000000_cam0_light0']
    with pytest.raises(ValueError):
        ds._get_nn_id({'cam': 'cam0', 'light': 'light0'})
    assert ds._get_nn_id({'cam': 'nope', 'light': 'light0'}) is None


def test_pipeline_batches_order_and_types(scene):
    cfg, _ = scene
    ds = _dataset(cfg, 'vali', n_map_parallel_calls=3, prefetch_buffer_size=2)
    pipe = ds.build_pipeline(pin_memory=False)
    batches = list(pipe)
    assert len(batches) == len(pipe) == 1
    b = batches[0]
    assert len(b) == 11 and b[0] == [b'trainvali_000000003_cam1_light1'] and isinstance(b[7][0], bytes)
    assert all(torch.is_tensor(t) and t.dtype == torch.float32 for i, t in enumerate(b) if i not in (0, 7))
    assert b[1].shape == (1, UVH, UVH, 3) and b[2].shape == (1, UVH, UVH, 1) and b[4].shape == (1, IMH, IMW, 2)
    # train: 3 examples, bs 2 -> a full and a short batch; same seed => same order, every example exactly once
    tr = _dataset(cfg, 'train', shuffle_buffer_size=2, n_map_parallel_calls=2)
    first = [i for bt in tr.build_pipeline(seed=5, pin_memory=False) for i in bt[0]]
    again = [i for bt in tr.build_pipeline(seed=5, pin_memory=False) for i in bt[0]]
    assert first == again and sorted(first) == sorted(f.encode() for f in tr.files)
    sizes = [len(bt[0]) for bt in tr.build_pipeline(seed=5, pin_memory=False)]
    assert sizes == [2, 1]
    # a second pass over the same pipe reshuffles (seed + pass index) but still covers every example
    pipe = tr.build_pipeline(seed=5, pin_memory=False)
    p1 = [i for bt in pipe for i in bt[0]]
    p2 = [i for bt in pipe for i in bt[0]]
    assert sorted(p1) == sorted(p2) == sorted(first)


def test_pipeline_take_shard_nobatch_filter(scene):
    cfg, _ = scene
    tr = _dataset(cfg, 'train')
    pipe = tr.build_pipeline(seed=1, pin_memory=False)
    assert len(list(pipe.take(1))) == 1 and len(list(pipe.take(-1))) == 2 and len(list(pipe.take(0))) == 0
    whole = list(tr.build_pipeline(seed=3, pin_memory=False))
    r0 = list(tr.build_pipeline(seed=3, pin_memory=False).shard(2, 0))
    r1 = list(tr.build_pipeline(seed=3, pin_memory=False).shard(2, 1))
    assert r0[0][0] + r1[0][0] == whole[0][0]                            # the two ranks split every global batch
    torch.testing.assert_close(torch.cat((r0[0][1], r1[0][1])), whole[0][1], rtol=0, atol=0)
    assert len(r0) == len(r1) == 1                                        # the 1-example last batch cannot be split: dropped
    with pytest.raises(ValueError):
        pipe.shard(2, 2)
    singles = list(tr.build_pipeline(seed=3, no_batch=True, pin_memory=False))
    assert len(singles) == 3 and singles[0][1].shape == (UVH, UVH, 3) and isinstance(singles[0][0], bytes)
    only = list(tr.build_pipeline(filter_predicate=lambda f: f.endswith('cam0_light0'), pin_memory=False))
    assert [i for bt in only for i in bt[0]] == [b'trainvali_000000000_cam0_light0']


def test_pipeline_cache_and_error_propagation(scene):
    cfg, _ = scene
    cfg['DEFAULT']['cache'] = 'True'
    tr = _dataset(cfg, 'train')
    calls = []
    orig = tr._process_example_precache
    tr._process_example_precache = lambda f: (calls.append(f), orig(f))[1]
    pipe = tr.build_pipeline(seed=0, pin_memory=False)
    list(pipe)
    list(pipe)
    assert sorted(calls) == sorted(tr.files)                             # second pass served from the cache
    cfg['DEFAULT']['cache'] = 'False'
    bad = _dataset(cfg, 'train')
    bad._process_example_precache = lambda f: (_ for _ in ()).throw(RuntimeError('decode failed: ' + f))
    with pytest.raises(RuntimeError, match='decode failed'):
        list(bad.build_pipeline(pin_memory=False))


def test_pipeline_batch_feeds_the_oracle_model(scene):
    """End of the chain: a pipeline batch has the layout the model's call() unpacks (nlt/models/nlt.py:91-92)."""
    cfg, _ = scene
    vali = _dataset(cfg, 'vali')
    (batch,) = list(vali.build_pipeline(pin_memory=False))
    id_, base, cvis, lvis, warp, rgb, rgb_camspc, nn_id, nn_base, nn_rgb, nn_rgb_camspc = batch
    x = torch.cat((base, cvis, lvis), dim=3)
    assert x.shape == (1, UVH, UVH, 5) and (nn_rgb - nn_base).shape == (1, UVH, UVH, 3)
    assert float(warp.min()) >= 0.0 and float(warp.max()) <= 1.0
    assert rgb_camspc.shape == (1, IMH, IMW, 3)


def test_trainvali_batch_source_real_and_synthetic(scene):
    """trainvali.batch_source: the on-disk dataset when its status JSON exists (sharded per rank, cycling over
    epochs), synthetic batches of the configured shape otherwise."""
    import types
    import trainvali
    cfg, _ = scene
    cfg['DEFAULT'].update(dict(dataset='nlt', no_batch='False', shuffle_buffer_size='4', prefetch_buffer_size='-1',
                               n_map_parallel_calls='2'))
    rank0 = types.SimpleNamespace(world=2, rank=0)
    rank1 = types.SimpleNamespace(world=2, rank=1)
    b0 = list(trainvali.batch_source(cfg, rank0, steps=3, device='cpu', seed=11))
    b1 = list(trainvali.batch_source(cfg, rank1, steps=3, device='cpu', seed=11))
    # one splittable batch per epoch (the short last batch is dropped by every rank): three epochs
    assert len(b0) == 3 and len(b1) == 3 and all(len(b[0]) == 1 for b in b0 + b1)
    assert b0[0][1].shape == (1, UVH, UVH, 3) and b0[0][1].dtype == torch.float32
    assert set(b0[0][0]).isdisjoint(b1[0][0])                             # the ranks see different examples
    cfg['DEFAULT']['data_root'] = cfg['DEFAULT']['data_root'] + '_nowhere'
    cfg['DEFAULT']['uvh'] = '16'
    cfg['DEFAULT']['imh'] = '16'
    syn = list(trainvali.batch_source(cfg, types.SimpleNamespace(world=1, rank=0), steps=2, device='cpu'))
    assert len(syn) == 2 and syn[0][1].shape == (2, 16, 16, 3) and not torch.equal(syn[0][1], syn[1][1])


def test_uint8_inputs_option_keeps_lossless_images_as_bytes(scene):
    """`uint8_inputs = True`: an 8-bit image that needs no resize is emitted as uint8 (the model divides by 255 on the
    device, bit-identical to float32(v / 255.0)); everything that is resized, and the warp, stays float32."""
    cfg, arrs = scene
    cfg_native = configparser.ConfigParser()
    cfg_native['DEFAULT'] = dict(cfg['DEFAULT'], uvh=str(UV_NATIVE), uvw=str(UV_NATIVE), imh=str(CAM_NATIVE[0]),
                                 imw=str(CAM_NATIVE[1]))
    id_ = 'trainvali_000000001_cam0_light1'
    ref = _dataset(cfg_native, 'train')._process_example_precache(id_)
    u8 = _dataset(cfg_native, 'train', uint8_inputs=True)._process_example_precache(id_)
    for i, (a, b) in enumerate(zip(ref, u8)):
        if i in (0, 7):
            assert a == b
        elif i == 4:
            assert b.dtype == np.float32 and np.array_equal(a, b)          # the warp
        else:
            assert b.dtype == np.uint8 and a.dtype == np.float32
            np.testing.assert_array_equal(a, b.astype(np.float32) / np.float32(255))
    # at a size that needs the cv2 resize nothing is kept as bytes
    rs = _dataset(cfg, 'train', uint8_inputs=True)._process_example_precache(id_)
    assert all(t.dtype == np.float32 for t in rs if isinstance(t, np.ndarray))
    # and the pipeline keeps the dtype through collation
    pipe = _dataset(cfg_native, 'train', uint8_inputs=True).build_pipeline(seed=0, pin_memory=False)
    b = next(iter(pipe))
    assert b[1].dtype == torch.uint8 and b[4].dtype == torch.float32 and b[1].shape[0] == 2
