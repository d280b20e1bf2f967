"""Host logic and ABI surface.  CPU only (no kernels are launched)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gen_feat_n_matches_reference_goldens():
    import numpy as np
    from util.net import gen_feat_n
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'gen_feat_n_reference.npz'))
    for key in gold.files:
        a, b, f = (int(v) for v in key.split('_'))
        assert gen_feat_n(a, b, f) == gold[key].tolist(), key
    with pytest.raises(AssertionError):
        gen_feat_n(16, 8)


def test_network_structure_matches_reference_wiring():
    from networks import convnet
    net = convnet.Network(16, 256, 2, 2, norm_type='None', act_type='leakyrelu', pool_type='None')
    assert len(net.layers) == 14
    assert net.is_contracting == [True] * 7 + [False] * 7
    assert net.spatsize_changes == [1] + [0.5] * 6 + [2] * 6 + [1]
    kinds = [[(c.kind, c.k, c.s, c.cout, c.act) for c in blk.convs] for blk in net.layers]
    assert kinds[0] == [('conv', 1, 1, 16, None)]
    assert kinds[1] == [('conv', 2, 2, 16, 'leakyrelu'), ('conv', 2, 1, 16, 'leakyrelu')]
    assert kinds[6] == [('conv', 2, 2, 256, 'leakyrelu'), ('conv', 2, 1, 256, 'leakyrelu')]
    assert kinds[7] == [('deconv', 2, 2, 128, 'leakyrelu'), ('deconv', 2, 1, 128, 'leakyrelu')]
    assert kinds[13] == [('conv', 1, 1, 3, None)]
    sss = convnet.Network(16, 1024, 2, 2, norm_type='None', act_type='relu', pool_type='None')
    assert len(sss.layers) == 18
    with pytest.raises(AssertionError):
        convnet.Network(16, 256, 2, 2, norm_type=None)          # str2none asserts on non-strings
    for bad in (dict(norm_type='instance'), dict(pool_type='max'), dict(act_type='gelu')):
        kw = dict(norm_type='None', act_type='relu', pool_type='None')
        kw.update(bad)
        with pytest.raises(NotImplementedError):
            convnet.Network(16, 256, 2, 2, **kw)


def test_loss_string_grammar():
    from models.base import Model
    p = Model._parse_loss_and_weight
    assert p('1e+0lpips') == ('lpips', 1.0)
    assert p('barron') == ('barron', 1.0)
    assert p('10l2') == ('l2', 10.0)
    assert p('1e+2l1') == ('l1', 100.0)
    assert p('.5ssim') == ('ssim', 0.5)


def test_model_registry_and_config_keys():
    import models
    from util import io as ioutil
    Model = models.get_model_class('nlt')
    for name in ('dragon_specular.ini', 'dragon_sss.ini'):
        cfg = ioutil.read_config(name)
        m = Model(cfg)
        m.register_trainable()
        assert m.trainable_registered and m.trainable_variables == []
        assert len(m.net['query'].layers) == (14 if 'specular' in name else 18)
    cfg = ioutil.make_config(loss='barron,1e+0lpips')
    with pytest.raises(NotImplementedError):
        Model(cfg)                                   # N1: not on the accelerated path yet
    with pytest.raises(NotImplementedError):
        Model(ioutil.make_config(loss='nope'))


def test_no_cpu_fallback():
    import models
    from util import io as ioutil, synth
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    m = models.get_model_class('nlt')(ioutil.make_config(uvh=64, uvw=64, imh=64, imw=64))
    m.register_trainable()
    import nlt_native as nat
    with pytest.raises(nat.NativeError):
        m(synth.make_batch(1, 64, 64), mode='test')


def test_same_pad():
    from engine import same_pad
    assert same_pad(8, 2, 1) == (0, 1) and same_pad(8, 2, 2) == (0, 0) and same_pad(8, 3, 2) == (0, 1)
    assert same_pad(7, 3, 2) == (1, 1) and same_pad(8, 3, 1) == (1, 1)


def test_abi_library_loads_and_exports_every_declared_symbol():
    import nlt_native as nat
    hdr = open(os.path.join(ROOT, 'include', 'nlt_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(nlt_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    assert declared == set(nat.exported_symbols())
    lib = nat.lib()          # loads the .so, binds every symbol (AttributeError if one is missing)
    raw = ctypes.CDLL(nat.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    assert b'sm_100a' in lib.nlt_version()
    # argument validation happens before any CUDA call
    d = nat.GConvDesc()
    assert lib.nlt_gconv_fwd(ctypes.byref(d), None, 0, 0.0, None, 0, None, None) == -1
    assert lib.nlt_gconv_wgrad_workspace_bytes(ctypes.byref(d)) == -1
    assert lib.nlt_l2_loss_workspace_bytes(4, 1000) > 0
    assert lib.nlt_uv2cam_fwd(None, None, None, None, 1, 1, 1, 1, 1, 1, None, None, None, None, None, None) == -1


def test_desc_struct_layout_matches_header():
    """ctypes mirror == C struct: compile a probe that prints sizeof/offsetof."""
    import nlt_native as nat
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "nlt_b200.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(nlt_gconv_desc), offsetof(nlt_gconv_desc, seg_ptr),
             offsetof(nlt_gconv_desc, seg_sub), offsetof(nlt_gconv_desc, seg_C), offsetof(nlt_gconv_desc, Cout),
             offsetof(nlt_gconv_desc, w), offsetof(nlt_gconv_desc, w_n_stride));
      return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, 'p.c'), 'w').write(src)
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), os.path.join(td, 'p.c'), '-o',
                               os.path.join(td, 'p')])
        out = subprocess.check_output([os.path.join(td, 'p')]).split()
    D = nat.GConvDesc
    want = [ctypes.sizeof(D), D.seg_ptr.offset, D.seg_sub.offset, D.seg_C.offset, D.Cout.offset, D.w.offset,
            D.w_n_stride.offset]
    assert [int(x) for x in out] == want


def test_synth_batch_properties():
    from util import synth
    b = synth.make_batch(2, 32, 32, seed=3, c_extra=4)
    assert b[1].shape == (2, 32, 32, 3) and b[2].shape == (2, 32, 32, 5) and b[4].shape == (2, 32, 32, 2)
    assert torch.equal(torch.round(b[1] * 255) / 255, b[1])
    frac_bg = float(((b[4] == 0).all(dim=3)).float().mean())
    assert 0.15 < frac_bg < 0.6
    assert torch.equal(b[4], b[4].half().float())
