"""LRU encoder layer: complex diagonal linear recurrence (Orvieto et al.) on the complex scan kernel
(kernels.complex_scan -> rorl_lru_scan_fwd/bwd).

Parameter names, initialiser and forward contract follow the reference layer
(ref: offpolicy_rnn/models/lru/lru.py:17-174; feed-forward tail :176-187).
"""
import numpy as np
import torch
import torch.nn as nn

from ..ensemble_linear_model import EnsembleLinear
from ..gilr.gilr import PositionWiseFeedForward
from ... import kernels as K


class LRULayer(nn.Module):
    def __init__(self, input_dim, output_dim, dropout=0.0, batch_first=True, use_ff=True, squash_inproj=False):
        super().__init__()
        assert batch_first, 'LRU only support batch_first==True'
        self.d_model = output_dim
        self.in_proj = EnsembleLinear(input_dim, self.d_model, num_ensemble=3, desire_ndim=4, bias=True)
        self.middle_proj = EnsembleLinear(self.d_model, self.d_model, num_ensemble=2, desire_ndim=4, bias=True)
        self.dropout = nn.Dropout(dropout)
        self.params_log = nn.Parameter(torch.vstack(self.initializer()), requires_grad=True)
        self.use_ff = use_ff
        self.squash_inproj = squash_inproj
        if use_ff:
            self.ff = PositionWiseFeedForward(self.d_model, dropout)

    def rnn_parameters(self):
        return self.parameters(recurse=True)

    def initializer(self):
        # ring initialisation |lambda| in [0.9, 0.999]  (ref: lru.py:48-67, arXiv:2303.06349 sec 3.2.2)
        r_min, r_max = 0.9, 0.999
        u1, u2 = torch.rand(self.d_model), torch.rand(self.d_model)
        nu_log = torch.log(-0.5 * torch.log(u1 * (r_max ** 2 - r_min ** 2) + r_min ** 2))
        theta_log = torch.log(u2 * torch.tensor(np.pi) * 2)
        lam = torch.exp(torch.complex(-torch.exp(nu_log), torch.exp(theta_log)))
        gamma_log = torch.log(torch.sqrt(1 - torch.abs(lam) ** 2))
        return nu_log, theta_log, gamma_log

    def forward(self, x, hidden=None, rnn_start=None, grad_detach=None):
        u = self.in_proj(x)                                    # [3, B, L, C]
        if self.squash_inproj:
            u = torch.tanh(u)
        nu, theta, gamma = torch.exp(self.params_log)
        mag = torch.exp(-nu)
        lam_re, lam_im = mag * torch.cos(theta), mag * torch.sin(theta)
        h0_re = h0_im = None
        if hidden is not None and not getattr(hidden, '_rorl_zero', False):
            h0_re, h0_im = hidden.transpose(0, 1).chunk(2, dim=-1)
        if grad_detach is None:
            # gamma scaling and the reset-gated decay lambda * (1 - start) are formed inside the scan kernel
            h_re, h_im = K.lru_fused_scan(u[0], u[1], lam_re, lam_im, gamma, rnn_start, h0_re, h0_im)
        else:
            v_re, v_im = gamma * u[0], gamma * u[1]
            keep = 1.0 if rnn_start is None else (1 - rnn_start)
            f_re = (lam_re * keep).expand_as(v_re)
            f_im = (lam_im * keep).expand_as(v_im)
            h_re, h_im = K.complex_scan(v_re, v_im, f_re, f_im, h0_re, h0_im, grad_detach)
        new_hidden = torch.cat((h_re[:, -1:, :], h_im[:, -1:, :]), dim=-1).transpose(0, 1)
        m = self.middle_proj(torch.stack((h_re, h_im), dim=0))
        out = m[0] - m[1] + u[2]
        if self.use_ff:
            out = self.ff(out)
        return out, new_hidden
