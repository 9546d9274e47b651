"""ORACLE (test infrastructure, not product code) -- CPU restatements of the reference's per-op
arithmetic on the update hot path, in plain PyTorch so autograd supplies the gradient oracle.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  Pinned against fixtures generated from the unmodified reference by
tests/golden/make_golden.py (see tests/test_oracle_golden.py).

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def gilr_scan(v: torch.Tensor, f: torch.Tensor, h0: torch.Tensor = None):
    """h_t = h_{t-1} * f_t + v_t * (1 - f_t); returns (h [B,L,C], h_last [B,1,C]).
    ref: offpolicy_rnn/models/gilr/scan_triton/real_rnn_tie_input_gate_cpu.py:4-14"""
    B, L, C = v.shape
    h = torch.zeros((B, C), dtype=v.dtype) if h0 is None else h0.reshape(B, C)
    out = []
    for t in range(L):
        h = h * f[:, t] + v[:, t] * (1 - f[:, t])
        out.append(h)
    return torch.stack(out, dim=1), h.unsqueeze(1)


def lru_scan(v_re, v_im, f_re, f_im, h_re=None, h_im=None, grad_detach=None):
    """complex h_t = f_t * h_{t-1} + v_t; returns (H_re, H_im).
    forward: ref offpolicy_rnn/models/lru/scan_triton/complex_rnn_cpu.py:4-26.
    grad_detach[b,t]=1 cuts the gradient flowing from step t back to step t-1's state, exactly the
    `grad_h *= (1 - grad_detach_t)` of the Triton backward (ref: complex_rnn.py:137-141)."""
    B, L, C = v_re.shape
    hr = torch.zeros((B, C), dtype=v_re.dtype) if h_re is None else h_re.reshape(B, C).detach()
    hi = torch.zeros((B, C), dtype=v_re.dtype) if h_im is None else h_im.reshape(B, C).detach()
    outs_r, outs_i = [], []
    for t in range(L):
        nr = hr * f_re[:, t] - hi * f_im[:, t] + v_re[:, t]
        ni = hr * f_im[:, t] + hi * f_re[:, t] + v_im[:, t]
        outs_r.append(nr)
        outs_i.append(ni)
        hr, hi = nr, ni
        if grad_detach is not None:
            # the gradient that step t+1.. sends into h_t is scaled by (1 - gd_t) ... the Triton kernel
            # applies the factor of step t to the carry arriving at step t (from t+1):
            gd = grad_detach.reshape(B, L)[:, t].unsqueeze(-1)
            hr = hr * (1 - gd) + (hr * gd).detach()
            hi = hi * (1 - gd) + (hi * gd).detach()
    return torch.stack(outs_r, 1), torch.stack(outs_i, 1)


def selective_scan(u, delta, A, B, C, start=None, D=None, z=None, delta_bias=None, delta_softplus=False,
                   return_last_state=False):
    """Mamba S6 scan with reset flag, reference layout u/delta/z/start [B,D,L], B/C [B,N,L], A [D,N].
    ref: offpolicy_rnn/models/smamba/mamba_ssm/ops/selective_scan_interface_new.py:96-166
    (bias+softplus :117-120, deltaA with reset :131-135, recurrence :147-160, D skip :161, gate :162-163).
    Restated step-by-step (no [B,D,L,N] materialisation) so it runs at L ~ 1000 on CPU."""
    u = u.float()
    delta = delta.float()
    if delta_bias is not None:
        delta = delta + delta_bias[..., None].float()
    if delta_softplus:
        delta = F.softplus(delta)
    Bsz, Dm, L = u.shape
    x = torch.zeros((Bsz, Dm, A.shape[1]), dtype=torch.float32)
    ys = []
    for t in range(L):
        dA = torch.exp(delta[:, :, t, None] * A[None])
        if start is not None:
            dA = dA * (1 - start[:, :, t, None])
        x = dA * x + (delta[:, :, t] * u[:, :, t])[..., None] * B[:, None, :, t]
        ys.append((x * C[:, None, :, t]).sum(-1))
    y = torch.stack(ys, dim=2)
    out = y if D is None else y + u * D[None, :, None]
    if z is not None:
        out = out * F.silu(z)
    return (out, x) if return_last_state else out


def s6_scan(u, delta, A, B, C, D, start, h0=None):
    """The s6 layer's CPU scan, token-major u/delta [B,L,D], B/C [B,L,N], start [B,L,1]; returns (y, final state).
    ref: offpolicy_rnn/models/s6/selective_scan/cpu_scan.py:6-61 (deltaA with reset :44-45, deltaB_u :46,
    initial state :52-53, recurrence :55-58, D skip :61).  Step-by-step, no [B,L,D,N] materialisation."""
    Bsz, L, Dm = u.shape
    x = torch.zeros((Bsz, Dm, A.shape[1]), dtype=torch.float32)
    if h0 is not None:
        x = x + h0
    keep = 1 - start.reshape(Bsz, L)
    ys = []
    for t in range(L):
        dA = torch.exp(delta[:, t, :, None] * A[None]) * keep[:, t, None, None]
        x = dA * x + (delta[:, t] * u[:, t])[..., None] * B[:, t, None, :]
        ys.append((x * C[:, t, None, :]).sum(-1))
    y = torch.stack(ys, dim=1)
    return y + u * D[None, None, :], x


def add_norm(x, weight, bias, residual=None, eps=1e-6, prenorm=False, is_rms=False):
    """residual add then LayerNorm / RMSNorm; prenorm=True also returns the sum.
    ref: offpolicy_rnn/models/smamba/mamba_ssm/ops/triton/layernorm_cpu.py:6-35"""
    x = x.float()
    if residual is not None:
        x = x + residual.float()
    if is_rms:
        rstd = 1 / torch.sqrt(x.square().mean(dim=-1, keepdim=True) + eps)
        out = x * rstd * weight
        if bias is not None:
            out = out + bias
    else:
        out = F.layer_norm(x, x.shape[-1:], weight=weight, bias=bias, eps=eps)
    return (out, x) if prenorm else out


def causal_conv1d_silu(x, weight, bias, mask=None):
    """x [B,D,L] (reference layout); depthwise causal conv (padding K-1, truncated to L) + SiLU with the
    mask applied to the conv input.  ref: offpolicy_rnn/models/smamba/mamba.py:75-83,207-212"""
    if mask is not None:
        x = mask * x
    L = x.shape[-1]
    K = weight.shape[-1]
    y = F.conv1d(x, weight, bias, padding=K - 1, groups=x.shape[1])[..., :L]
    return F.silu(y)


def ensemble_linear(x, weight, bias, desire_ndim=None):
    """ref: offpolicy_rnn/models/ensemble_linear_model.py:29-60"""
    E = weight.shape[0]
    if x.dim() == 2:
        x = torch.einsum('ij,bjk->bik', x, weight)
    elif x.dim() == 3:
        if (desire_ndim is None or desire_ndim == 3) and x.shape[0] == E:
            x = torch.einsum('bij,bjk->bik', x, weight)
        else:
            x = torch.einsum('cij,bjk->bcik', x, weight)
    elif x.dim() == 4:
        if (desire_ndim is None or desire_ndim == 4) and x.shape[0] == E:
            x = torch.einsum('cbij,cjk->cbik', x, weight)
        else:
            x = torch.einsum('cdij,bjk->bcdik', x, weight)
    elif x.dim() == 5:
        x = torch.einsum('bcdij,bjk->bcdik', x, weight)
    if bias is not None:
        b = bias
        if x.dim() == 4:
            b = b.unsqueeze(1)
        elif x.dim() == 5:
            b = b.unsqueeze(1).unsqueeze(1)
        x = x + b
    return x


def tanh_gaussian(mean_raw, logstd, noise):
    """ref: offpolicy_rnn/policy_value_models/contextual_sac_policy_single_head.py:109-123"""
    logstd = torch.clamp(logstd, -20.0, 2.0)
    std = logstd.exp()
    sample = mean_raw + noise * std
    logp = (-0.5 * noise.pow(2) - (logstd + 0.5 * math.log(2 * math.pi))).sum(-1, keepdim=True)
    logp = logp - (2 * (-sample - F.softplus(-2 * sample) + math.log(2))).sum(-1, keepdim=True)
    return torch.tanh(mean_raw), torch.tanh(sample), logp


class QValueGuard:
    """ref: offpolicy_rnn/utility/q_value_guard.py:4-45 (constructed with guard_min=guard_max=True)"""

    def __init__(self, decay_ratio=1.0):
        self.min, self.max, self.init, self.decay = 1000000, -1000000, True, decay_ratio

    def clamp(self, value):
        if self.init:
            self.min, self.max, self.init = value.min().item(), value.max().item(), False
        return value.clamp(min=self.min, max=self.max)

    def update(self, value):
        vmin, vmax = value.min().item(), value.max().item()
        self.min, self.max = min(self.min, vmin), max(self.max, vmax)
        if self.decay < 1:
            self.min = self.decay * self.min + (1 - self.decay) * vmin
            self.max = self.decay * self.max + (1 - self.decay) * vmax
