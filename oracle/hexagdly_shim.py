"""Drop-in stand-in for the un-vendored ``hexagdly`` package (test infrastructure only).

Lets /root/reference/gridnext/gridnet_models.py import unmodified in the build container
(``sys.modules['hexagdly'] = oracle.hexagdly_shim``) so golden vectors can be generated from
the reference's own module code.  Interface restated from upstream HexagDLy:
``Conv2d(in_channels, out_channels, kernel_size=1, stride=1, bias=True, debug=False)`` with
parameters ``kernel0..kernel{k}`` and ``bias_tensor`` initialised U(-1/sqrt(n), 1/sqrt(n)),
n = in_channels * (1 + 3k(k+1)).  PARITY UNPINNED (see oracle/hexconv_ref.py).
"""
import math
import torch
import torch.nn as nn
from .hexconv_ref import hexconv_hexagdly, kernel_shapes, n_taps


class Conv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, bias=True, debug=False):
        super().__init__()
        if stride != 1:
            raise NotImplementedError("oracle shim restates stride 1 only (all the hot path uses)")
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.hexbase_size = kernel_size
        self.hexbase_stride = stride
        self.debug = debug
        self.bias = bias
        for i, shp in enumerate(kernel_shapes(in_channels, out_channels, kernel_size)):
            setattr(self, 'kernel' + str(i), nn.Parameter(torch.empty(shp)))
        if bias:
            self.bias_tensor = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter('bias_tensor', None)
        self.init_parameters(debug)

    def init_parameters(self, debug):
        if debug:
            for p in self.parameters():
                p.data.fill_(1.0)
            if self.bias_tensor is not None:
                self.bias_tensor.data.fill_(0.0)
            return
        stdv = 1.0 / math.sqrt(self.in_channels * n_taps(self.hexbase_size))
        for p in self.parameters():
            p.data.uniform_(-stdv, stdv)

    def forward(self, x):
        ks = [getattr(self, 'kernel' + str(i)) for i in range(self.hexbase_size + 1)]
        return hexconv_hexagdly(x, ks, self.bias_tensor)

    def __repr__(self):
        return 'Conv2d({}, {}, kernel_size={}, stride={})'.format(
            self.in_channels, self.out_channels, self.hexbase_size, self.hexbase_stride)
