"""`conv1d_*` encoder layer: depthwise causal convolution over time (no activation) + the feed-forward tail, with the
window of the last d_conv - 1 masked inputs as its hidden state.

Parameter names and forward contract follow the reference (ref: offpolicy_rnn/models/conv1d/conv1d.py:5-52): the
conv takes an explicit left state and padding 0 (:26-35), `mask` zeroes the conv input, `rnn_start` is not looked at,
the hidden comes in as [1, B, (K-1)*C] and goes out batch-first as [B, 1, (K-1)*C].  The convolution runs on
rorl_conv1d_fwd / _bwd (csrc/conv1d.cu) with the activation switched off.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import kernels as K
from ..gilr.gilr import PositionWiseFeedForward as _FF


class PositionWiseFeedForward(_FF):
    """ref: conv1d.py:57-68 (dropout 0.0, LayerNorm at torch's default eps)"""

    def __init__(self, d_model, dropout=0.0):
        super().__init__(d_model, dropout, eps=1e-5)


class Conv1d(nn.Module):
    def __init__(self, in_channels, out_channels, d_conv=4, bias=True, ff=True):
        super().__init__()
        assert in_channels == out_channels
        self.in_channels, self.out_channels, self.d_conv = in_channels, out_channels, d_conv
        self.conv1d = nn.Conv1d(in_channels, out_channels, bias=bias, kernel_size=d_conv, groups=in_channels, padding=0)
        self.desired_hidden_dim = in_channels * (d_conv - 1)
        self.use_ff = ff
        if ff:
            self.ff = PositionWiseFeedForward(out_channels, 0.0)

    def forward(self, x, hidden=None, mask=None):
        B, L, C = x.shape
        Kc = self.d_conv
        carried = hidden is not None and not getattr(hidden, '_rorl_zero', False)
        xm = x if mask is None else x * mask
        if carried:
            x_in = torch.cat((hidden.reshape(B, Kc - 1, C), xm), dim=1)
            y = K.causal_conv1d(x_in, self.conv1d.weight, self.conv1d.bias, None)[:, Kc - 1:, :]
        else:
            x_in = xm
            y = K.causal_conv1d(x, self.conv1d.weight, self.conv1d.bias, mask)
        tail = x_in[:, -(Kc - 1):, :] if Kc > 1 else x_in[:, :0]
        if tail.shape[1] < Kc - 1:
            tail = F.pad(tail, (0, 0, Kc - 1 - tail.shape[1], 0))
        new_hidden = tail.detach().reshape(B, 1, -1)
        if self.use_ff:
            y = self.ff(y)
        return y, new_hidden
