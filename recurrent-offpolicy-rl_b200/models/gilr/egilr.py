"""`egilr-E` encoder: E independent GILR layers evaluated side by side (one encoder per ensemble member), output
[E, B, L, C].  Parameters, hidden layout ([1, E*B, C] out) and forward contract follow the reference
(ref: offpolicy_rnn/models/gilr/egilr.py:15-98); the recurrence runs on the fused GILR scan kernel over the [E*B] rows."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import kernels as K
from ..ensemble_linear_model import EnsembleLinear
from ..multi_ensemble_linear_model import MultiEnsembleLinear


class EnsemblePositionWiseFeedForward(nn.Module):
    """w2(GELU(w1 x)) + x, then LayerNorm over the joint (ensemble, feature) axes (ref: egilr.py:84-98)."""

    def __init__(self, d_model, n_ensemble, dropout=0.1, desire_ndim=4):
        super().__init__()
        self.w_1 = EnsembleLinear(d_model, d_model, n_ensemble, desire_ndim=desire_ndim)
        self.w_2 = EnsembleLinear(d_model, d_model, n_ensemble, desire_ndim=desire_ndim)
        self.activation = nn.GELU()
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm([n_ensemble, d_model])

    def forward(self, x):
        y = self.dropout(self.activation(self.w_1(x)))
        y = self.dropout(self.w_2(y)) + x
        return self.layer_norm(y.transpose(0, -2)).transpose(0, -2)


class EnsembleGILRLayer(nn.Module):
    def __init__(self, input_dim, output_dim, num_ensemble, factor=1, dropout=0.0, use_ff=True, batch_first=True):
        super().__init__()
        assert batch_first
        self.d_model, self.num_ensemble = output_dim, num_ensemble
        self.in_proj = MultiEnsembleLinear(input_dim, self.d_model * factor, num_ensemble, 2, desire_ndim=4)
        self.out_proj = EnsembleLinear(self.d_model * factor, self.d_model * factor, num_ensemble)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm([num_ensemble, factor * self.d_model])      # constructed but unused, as in the reference
        self.swish = nn.SiLU()
        self.use_ff = use_ff
        if use_ff:
            self.ff = EnsemblePositionWiseFeedForward(self.d_model, num_ensemble, dropout, desire_ndim=4)

    def rnn_parameters(self):
        return list(self.parameters(True))

    def forward(self, x, hidden=None, rnn_start=None):
        u = self.in_proj(x)                                                      # [2, E, B, L, C]
        E = self.num_ensemble
        Bsz, L, C = u.shape[2], u.shape[3], u.shape[4]
        u_v, u_f = u[0].reshape(E * Bsz, L, C), u[1].reshape(E * Bsz, L, C)
        start = None
        if rnn_start is not None:
            start = (rnn_start if rnn_start.dim() == 4 else rnn_start.unsqueeze(0).expand(E, *rnn_start.shape)).reshape(E * Bsz, L, 1)
        h = K.gilr_fused_scan(u_v, u_f, start)
        if hidden is not None and not getattr(hidden, '_rorl_zero', False) and bool((hidden != 0).any()):
            f = torch.sigmoid(u_f)
            if start is not None:
                f = f * (1 - start)
            h = h + torch.cumprod(f, dim=1) * hidden.transpose(0, 1)            # carried state [1, E*B, C]
        new_hidden = h[:, -1:, :].transpose(0, 1)
        out = self.out_proj(h.reshape(E, Bsz, L, C))
        if self.use_ff:
            out = self.ff(out)
        return out, new_hidden
