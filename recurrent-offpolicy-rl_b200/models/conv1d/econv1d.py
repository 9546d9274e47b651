"""`econv1d[_K]-E` encoder: E independent depthwise causal convolutions (one per ensemble member), output [E, B, L, C].
Parameters (`conv1d.weight [E*C, 1, K]`), hidden layout ([B, 1, E*(K-1)*C] out) and forward contract follow the
reference (ref: offpolicy_rnn/models/conv1d/econv1d.py:3-88): a shared [B, L, C] input is repeated per member, `mask`
zeroes the conv input, no activation and no feed-forward tail.  Runs on rorl_conv1d_fwd / _bwd with the members as extra
channels of the token-major layout."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import kernels as K


class EConv1d(nn.Module):
    def __init__(self, in_channels, out_channels, num_ensemble, d_conv=4, bias=True):
        super().__init__()
        assert in_channels == out_channels
        self.in_channels, self.out_channels, self.num_ensemble, self.d_conv = in_channels, out_channels, num_ensemble, d_conv
        self.conv1d = nn.Conv1d(in_channels * num_ensemble, out_channels * num_ensemble, bias=bias, kernel_size=d_conv,
                                groups=in_channels * num_ensemble, padding=0)
        self.desired_hidden_dim = in_channels * (d_conv - 1) * num_ensemble
        self.desire_ndim = 4

    def forward(self, x, hidden=None, mask=None):
        E, C, Kc = self.num_ensemble, self.in_channels, self.d_conv
        if x.dim() == 3 and self.desire_ndim == 4:
            x = x.unsqueeze(0).repeat_interleave(E, dim=0)
        else:
            assert x.dim() == 4
        _, B, L, _ = x.shape
        xt = x.permute(1, 2, 0, 3).reshape(B, L, E * C)                       # token-major, channel index = e * C + c
        carried = hidden is not None and not getattr(hidden, '_rorl_zero', False)
        xm = xt if mask is None else xt * mask
        if carried:
            h = hidden.reshape(B, E, Kc - 1, C).permute(0, 2, 1, 3).reshape(B, Kc - 1, E * C)
            x_in = torch.cat((h, xm), dim=1)
            y = K.causal_conv1d(x_in, self.conv1d.weight, self.conv1d.bias, None)[:, Kc - 1:, :]
        else:
            x_in = xm
            y = K.causal_conv1d(xt.contiguous(), self.conv1d.weight, self.conv1d.bias, mask)
        tail = x_in[:, -(Kc - 1):, :] if Kc > 1 else x_in[:, :0]
        if tail.shape[1] < Kc - 1:
            tail = F.pad(tail, (0, 0, Kc - 1 - tail.shape[1], 0))
        new_hidden = tail.detach().reshape(B, Kc - 1, E, C).permute(0, 2, 1, 3).reshape(B, 1, -1)
        return y.reshape(B, L, E, C).permute(2, 0, 1, 3), new_hidden
