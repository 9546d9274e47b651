"""GILR encoder layer: gated linear recurrence h = f*h + (1-f)*v with the tanh / sigmoid / reset gating
fused into the scan kernel (kernels.gilr_fused_scan -> rorl_gilr_fused_fwd/bwd).

Parameter names and forward contract follow the reference layer
(ref: offpolicy_rnn/models/gilr/gilr.py:13-67; feed-forward tail :70-81).  Differences, by design:
no `torch.all(hidden == 0)` host sync (ref :57) -- a non-zero carried state is folded in exactly, by
linearity, as a rank-1 correction; any hidden width C % 4 == 0 (the reference needs C % 256 == 0).
"""
import torch
import torch.nn as nn

from ..ensemble_linear_model import EnsembleLinear
from ... import kernels as K
from ..linear import Linear


class PositionWiseFeedForward(nn.Module):
    def __init__(self, d_model, dropout=0.1, eps=1e-5):
        super().__init__()
        self.w_1 = Linear(d_model, d_model)
        self.w_2 = Linear(d_model, d_model)
        self.activation = nn.GELU()
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(d_model, eps=eps)

    def forward(self, x):
        y = self.dropout(self.w_2(self.dropout(self.activation(self.w_1(x)))))
        if not x.is_cuda:
            return self.layer_norm(y + x)
        # residual add + LayerNorm in one kernel (ATen's LayerNorm backward ran at 240 us per call on [32, 1002, 256])
        return K.layer_norm_fn(y, self.layer_norm.weight, self.layer_norm.bias, residual=x, eps=self.layer_norm.eps, prenorm=False)


class GILRLayer(nn.Module):
    def __init__(self, input_dim, output_dim, factor=1, dropout=0.0, use_ff=True, batch_first=True):
        super().__init__()
        assert batch_first
        self.d_model = output_dim
        self.in_proj = EnsembleLinear(input_dim, self.d_model * factor, 2, desire_ndim=4)
        self.out_proj = Linear(self.d_model * factor, self.d_model * factor)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(factor * self.d_model)   # constructed but unused, as in the reference
        self.use_ff = use_ff
        if use_ff:
            self.ff = PositionWiseFeedForward(self.d_model, dropout)

    def rnn_parameters(self):
        return list(self.parameters(True))

    def forward(self, x, hidden=None, rnn_start=None):
        u = self.in_proj(x)                                   # [2, B, L, C]
        h = K.gilr_fused_scan(u[0], u[1], rnn_start)
        if hidden is not None and not getattr(hidden, '_rorl_zero', False):
            # carried state: h_t += (prod_{s<=t} f_s) * h_prev.  Zero at update time (make_init_state).
            h = h + _carry_correction(u[1], rnn_start, hidden)
        new_hidden = h[:, -1:, :].transpose(0, 1)
        out = self.out_proj(h)
        if self.use_ff:
            out = self.ff(out)
        return out, new_hidden


def _carry_correction(u_f, rnn_start, hidden):
    hprev = hidden.transpose(0, 1)                            # [B, 1, C]
    f = torch.sigmoid(u_f)
    if rnn_start is not None:
        f = f * (1 - rnn_start)
    return torch.cumprod(f, dim=1) * hprev
