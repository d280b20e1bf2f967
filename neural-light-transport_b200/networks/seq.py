"""Mirror of nlt/networks/seq.py:27-41 (simple sequential flow)."""
import torch

from .base import Network as BaseNetwork


class Network(BaseNetwork):
    def build(self, input_shape):
        """input_shape: (N, H, W, C) like Keras' Sequential.build."""
        cin = int(input_shape[-1])
        dev = torch.device('cuda', torch.cuda.current_device())
        for layer in self.layers:
            cin = layer.build(cin, dev)
        for layer in self.layers:
            assert layer.built, "Some layers not built"

    def __call__(self, tensor):
        x = tensor
        y = None
        for layer in self.layers:
            y = layer(x)
            x = y
        return y
