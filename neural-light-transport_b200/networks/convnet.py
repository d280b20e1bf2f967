"""UV-space encoder/decoder layer list -- mirror of nlt/networks/convnet.py:30-90.

Same constructor, same attributes (.layers, .is_contracting,
.spatsize_changes); each entry of .layers is a callable Block (x -> y on NHWC
CUDA tensors), which is what nlt/nlt_test.py:111-114 and
nlt/models/nlt.py:147-198 iterate over.
"""
import numpy as np
import torch

from engine import Act, Seg
from util import net as netutil
from .seq import Network as BaseNetwork
from .elements import conv, norm, act, pool, deconv


class Block:
    """One entry of Network.layers: 1x1 conv, down block or up block, run as
    fused conv(+bias+act) CUDA ops."""

    def __init__(self, convs):
        self.convs = convs

    @property
    def built(self):
        return all(c.built for c in self.convs)

    def build(self, cin, device, generator=None):
        for c in self.convs:
            if not c.built:
                c.build(cin, device, generator)
            cin = c.cout
        return cin

    def forward_segs(self, segs, tape=None):
        """segs: list of engine.Seg (virtual concat) -> engine.Act"""
        y = self.convs[0].forward(segs, tape)
        for c in self.convs[1:]:
            y = c.forward([Seg(y)], tape)
        return y

    def __call__(self, x):
        if isinstance(x, Act):
            return self.forward_segs([Seg(x)])
        return self.forward_segs([Seg(Act(x.contiguous()))]).t


class Network(BaseNetwork):
    def __init__(
            self, depth0, depth, kernel, stride, norm_type=None,
            act_type='relu', pool_type=None):
        super().__init__()
        norm_type = self.str2none(norm_type)
        pool_type = self.str2none(pool_type)
        norm(norm_type)   # raises NotImplementedError for unsupported kinds
        pool(pool_type)
        a = act(act_type)
        n_feat = netutil.gen_feat_n(depth0, depth)
        prev_n = 0
        self.is_contracting, self.spatsize_changes = [], []
        # 1x1 conv to generate an original-res. feature map (convnet.py:44)
        self.layers.append(Block([conv(1, n_feat[0], stride=1)]))
        self.is_contracting.append(True)
        self.spatsize_changes.append(1)
        for n in n_feat[:-1]:
            if n >= prev_n:   # so 64 -> 64 is considered "contracting" (:49)
                c1, c2 = conv(kernel, n, stride=stride), conv(kernel, n, stride=1)
                c1.act = c2.act = a
                self.layers.append(Block([c1, c2]))
                self.is_contracting.append(True)
                self.spatsize_changes.append(1 / stride)
            else:
                d1, d2 = deconv(kernel, n, stride=stride), deconv(kernel, n, stride=1)
                d1.act = d2.act = a
                self.layers.append(Block([d1, d2]))
                self.is_contracting.append(False)
                self.spatsize_changes.append(stride)
            prev_n = n
        # final 1x1 conv (convnet.py:85)
        self.layers.append(Block([conv(1, n_feat[-1], stride=1)]))
        self.is_contracting.append(False)
        self.spatsize_changes.append(1)
        spatsizes = np.cumprod(self.spatsize_changes)
        assert spatsizes[-1] == 1, \
            "Resolution doesn't return to the original value"

    def conv_layers(self):
        out = []
        for blk in self.layers:
            out += blk.convs
        return out
