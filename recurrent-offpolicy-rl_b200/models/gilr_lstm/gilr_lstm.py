"""`gilr_lstm` encoder layer: a GILR scan feeding an LSTM-style gated linear recurrence.

    u = in_proj(x) [2, B, L, C];  v = scan(tanh(u0), sigmoid(u1) * (1 - start))            (first recurrence)
    g = middle_proj(v) [4, B, L, C];  f, i, o = sigmoid(g0..2), z = tanh(g3)
    c = scan(i * z, f * (1 - start));  out = out_proj(c * o)                                 (second recurrence)
with scan(v, f): h_t = f_t h_{t-1} + (1 - f_t) v_t.  Hidden = [first | second] recurrence state, 2C wide.

Parameter names and forward contract follow the reference (ref: offpolicy_rnn/models/gilr_lstm/gilr_lstm.py:13-75),
minus its `torch.all(hidden == 0)` host syncs (:50,62): a carried state is folded in by linearity.  Both scans run on
the gilr kernels (csrc/scan_real.cu): the first one fused with its gating, the second through the plain tied-gate scan.
"""
import torch
import torch.nn as nn

from ... import kernels as K
from ..ensemble_linear_model import EnsembleLinear
from ..linear import Linear


class GILRLSTMLayer(nn.Module):
    def __init__(self, input_dim, output_dim, factor=1, dropout=0.2, batch_first=True):
        super().__init__()
        assert batch_first
        self.d_model = output_dim
        self.in_proj = EnsembleLinear(input_dim, self.d_model * factor, 2, desire_ndim=4)
        self.middle_proj = EnsembleLinear(self.d_model * factor, self.d_model * factor, 4, desire_ndim=4)
        self.out_proj = Linear(self.d_model * factor, self.d_model * factor)
        self.layer_norm = nn.LayerNorm(factor * self.d_model)       # constructed but unused, as in the reference
        self.swish = nn.SiLU()

    def rnn_parameters(self):
        return list(self.parameters(True))

    @staticmethod
    def _carry(f, h_prev):
        """h_t += (prod_{s<=t} f_s) h_prev for a state carried into the call (zero at update time)."""
        return torch.cumprod(f, dim=1) * h_prev

    def forward(self, x, hidden=None, rnn_start=None):
        u = self.in_proj(x)                                            # [2, B, L, C]
        carried = hidden is not None and not getattr(hidden, '_rorl_zero', False)
        keep = None if rnn_start is None else (1 - rnn_start)
        v = K.gilr_fused_scan(u[0], u[1], rnn_start)
        if carried:
            h_pre, h_mid = torch.chunk(hidden.transpose(0, 1), 2, dim=-1)          # [B, 1, C] each
            f1 = torch.sigmoid(u[1])
            v = v + self._carry(f1 if keep is None else f1 * keep, h_pre)
        g = self.middle_proj(v)                                        # [4, B, L, C]
        f = torch.sigmoid(g[0])
        if keep is not None:
            f = f * keep
        iz = torch.sigmoid(g[1]) * torch.tanh(g[3])
        c = K.real_scan_tie_input_gate(iz, f)
        if carried:
            c = c + self._carry(f, h_mid)
        out = self.out_proj(c * torch.sigmoid(g[2]))
        new_hidden = torch.cat((v[:, -1:, :], c[:, -1:, :]), dim=-1).transpose(0, 1)
        return out, new_hidden
