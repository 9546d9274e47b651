"""`elru-E` (and `egilr_lstm-E`, which the reference maps to the same class, ref: rnn_base.py:110-113) encoder: E
independent LRU layers side by side, output [E, B, L, C].  Parameters (`params_log [3, E, C]`), hidden layout
([1, B, 2*C*E]: the E real parts then the E imaginary parts) and forward contract follow the reference
(ref: offpolicy_rnn/models/lru/elru.py:13-229); the recurrence runs on the complex scan kernel over the [E*B] rows."""
import numpy as np
import torch
import torch.nn as nn

from ... import kernels as K
from ..gilr.egilr import EnsemblePositionWiseFeedForward
from ..multi_ensemble_linear_model import MultiEnsembleLinear


class EnsembleLRULayer(nn.Module):
    def __init__(self, input_dim, output_dim, num_ensemble, dropout=0.0, batch_first=True, use_ff=True, squash_inproj=False):
        super().__init__()
        assert batch_first, 'LRU only support batch_first==True'
        self.d_model, self.num_ensemble = output_dim, num_ensemble
        self.in_proj = MultiEnsembleLinear(input_dim, self.d_model, num_ensemble, 3, desire_ndim=4, bias=True)
        self.middle_proj = MultiEnsembleLinear(input_dim, self.d_model, num_ensemble, 2, desire_ndim=4, bias=True)
        self.dropout = nn.Dropout(dropout)
        self.params_log = nn.Parameter(torch.stack(self.initializer(num_ensemble), dim=0), requires_grad=True)
        self.use_ff, self.squash_inproj = use_ff, squash_inproj
        if use_ff:
            self.ff = EnsemblePositionWiseFeedForward(self.d_model, num_ensemble, dropout, desire_ndim=4)

    def rnn_parameters(self):
        return self.parameters(recurse=True)

    def initializer(self, num_ensemble):
        r_min, r_max = 0.9, 0.999                                  # ring initialisation, ref: elru.py:53-62
        u1, u2 = torch.rand((num_ensemble, self.d_model)), torch.rand((num_ensemble, self.d_model))
        nu_log = torch.log(-0.5 * torch.log(u1 * (r_max ** 2 - r_min ** 2) + r_min ** 2))
        theta_log = torch.log(u2 * torch.tensor(np.pi) * 2)
        lam = torch.exp(torch.complex(-torch.exp(nu_log), torch.exp(theta_log)))
        gamma_log = torch.log(torch.sqrt(1 - torch.abs(lam) ** 2))
        return nu_log, theta_log, gamma_log

    def forward(self, x, hidden=None, rnn_start=None, grad_detach=None):
        u = self.in_proj(x)                                        # [3, E, B, L, C]
        if self.squash_inproj:
            u = torch.tanh(u)
        E = self.num_ensemble
        Bsz, L, C = u.shape[2], u.shape[3], u.shape[4]
        nu, theta, gamma = torch.exp(self.params_log)              # [E, C] each
        mag = torch.exp(-nu)
        lam_re, lam_im = mag * torch.cos(theta), mag * torch.sin(theta)
        g = gamma[:, None, None, :]
        v_re, v_im = (g * u[0]).reshape(E * Bsz, L, C), (g * u[1]).reshape(E * Bsz, L, C)
        keep = 1.0
        if rnn_start is not None:
            keep = 1 - (rnn_start if rnn_start.dim() == 4 else rnn_start.unsqueeze(0))
        f_re = (lam_re[:, None, None, :] * keep).expand(E, Bsz, L, C).reshape(E * Bsz, L, C)
        f_im = (lam_im[:, None, None, :] * keep).expand(E, Bsz, L, C).reshape(E * Bsz, L, C)
        h0_re = h0_im = None
        if hidden is not None and not getattr(hidden, '_rorl_zero', False):
            items = hidden.transpose(0, 1).chunk(2 * E, dim=-1)    # 2E pieces of [B, 1, C]: E real parts, then E imaginary parts
            h0_re = torch.cat([t.unsqueeze(0) for t in items[:E]], dim=0).reshape(E * Bsz, 1, C)
            h0_im = torch.cat([t.unsqueeze(0) for t in items[E:]], dim=0).reshape(E * Bsz, 1, C)
        gd = None
        if grad_detach is not None:
            gd = grad_detach
            if gd.shape[0] < E * Bsz:
                gd = gd.unsqueeze(0).repeat_interleave(E, dim=0).reshape(E * Bsz, gd.shape[-2], gd.shape[-1])
        h_re, h_im = K.complex_scan(v_re, v_im, f_re, f_im, h0_re, h0_im, gd)
        h_re, h_im = h_re.reshape(E, Bsz, L, C), h_im.reshape(E, Bsz, L, C)
        last_re = torch.cat([h_re[i, :, -1:, :] for i in range(E)], dim=-1)      # [B, 1, C*E]
        last_im = torch.cat([h_im[i, :, -1:, :] for i in range(E)], dim=-1)
        new_hidden = torch.cat((last_re, last_im), dim=-1).transpose(0, 1)
        m = self.middle_proj(torch.stack((h_re, h_im), dim=0))    # [2, E, B, L, C]
        out = m[0] - m[1] + u[2]
        if self.use_ff:
            out = self.ff(out)
        return out, new_hidden
