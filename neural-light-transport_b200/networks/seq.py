"""Sequential network: `build(input_shape)` then `net(x)` runs the layers in order (interface of
nlt/networks/seq.py:27-41).  Layers are engine-backed blocks whose `build(cin, device)` returns their output
channel count, so building is a fold over the channel dimension of `input_shape` (NHWC)."""
import functools

import torch

from .base import Network as BaseNetwork


class Network(BaseNetwork):
    def build(self, input_shape):
        device = torch.device('cuda', torch.cuda.current_device())
        functools.reduce(lambda cin, layer: layer.build(cin, device), self.layers, int(input_shape[-1]))
        assert all(layer.built for layer in self.layers), "Some layers not built"

    def __call__(self, tensor):
        out = None
        for layer in self.layers:
            out = tensor = layer(tensor)
        return out
