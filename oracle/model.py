"""ORACLE (test infrastructure) -- functional CPU restatement of the reference's model stack on the
update path: layer registry -> encoder layers -> policy / value heads, driven by state dicts whose key
names are the reference's own.  Plain PyTorch ops + autograd.

ref: offpolicy_rnn/models/rnn_base.py (RNNBase), contextual_model.py (ContextualModel),
     policy_value_models/contextual_sac_policy_single_head.py, contextual_sac_value.py,
     contextual_td3_policy.py, models/smamba/mamba.py (GPU path = forward_sequential),
     models/gilr/gilr.py, models/lru/lru.py, models/s6/mamba.py (+ s6/selective_scan/cpu_scan.py), torch.nn.GRU.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import ops

ACT = {'tanh': torch.tanh, 'relu': F.relu, 'sigmoid': torch.sigmoid, 'leaky_relu': F.leaky_relu,
       'linear': lambda x: x, 'elu': F.elu, 'gelu': F.gelu}   # ref: rnn_base.py:45-53


def parse_smamba(layer_id: str):
    """ref: rnn_base.py:137-163"""
    cfg = dict(d_state=16, d_conv=4, blocks=2, rms=True, ff=False)
    for tok in layer_id.split('_')[1:]:
        if tok.startswith('s'):
            cfg['d_state'] = int(tok[1:])
        elif tok.startswith('c'):
            cfg['d_conv'] = int(tok[1:])
        elif tok.startswith('b'):
            cfg['blocks'] = int(tok[1:])
        elif tok.startswith('n'):
            cfg['rms'] = tok[1:] != 'ln'
        elif tok.startswith('f'):
            cfg['ff'] = tok[1:] == 'f'
    return cfg


def is_rnn(t: str) -> bool:
    return not (t == 'fc' or t.startswith('efc'))


class Side:
    """RNNHidden side-band (ref: offpolicy_rnn/models/RNNHidden.py:36-62)."""

    def __init__(self, rnn_start=None, mask=None, attention_concat_mask=None, grad_detach=None, h0=None):
        self.rnn_start, self.mask, self.attention_concat_mask, self.grad_detach = rnn_start, mask, attention_concat_mask, grad_detach
        self.h0 = h0 or {}


def ff_block(p: Dict[str, torch.Tensor], pre: str, x, eps=1e-5):
    """PositionWiseFeedForward: LN(w2(gelu(w1 x)) + x)  ref: gilr/gilr.py:70-81, lru/lru.py:176-187,
    smamba/mamba.py:528-539"""
    y = F.linear(F.gelu(F.linear(x, p[pre + 'w_1.weight'], p[pre + 'w_1.bias'])), p[pre + 'w_2.weight'], p[pre + 'w_2.bias'])
    return F.layer_norm(y + x, x.shape[-1:], p[pre + 'layer_norm.weight'], p[pre + 'layer_norm.bias'], eps)


def gilr_layer(p, pre, x, side: Side):
    """ref: offpolicy_rnn/models/gilr/gilr.py:44-67"""
    u = ops.ensemble_linear(x, p[pre + 'in_proj.weight'], p[pre + 'in_proj.bias'], desire_ndim=4)
    v, f = torch.tanh(u[0]), torch.sigmoid(u[1])
    if side.rnn_start is not None:
        f = f * (1 - side.rnn_start)
    h, _ = ops.gilr_scan(v, f)
    out = F.linear(h, p[pre + 'out_proj.weight'], p[pre + 'out_proj.bias'])
    return ff_block(p, pre + 'ff.', out)


def lru_layer(p, pre, x, side: Side):
    """ref: offpolicy_rnn/models/lru/lru.py:70-174"""
    u = ops.ensemble_linear(x, p[pre + 'in_proj.weight'], p[pre + 'in_proj.bias'], desire_ndim=4)
    params = torch.exp(p[pre + 'params_log'])
    nu, theta, gamma = params[0], params[1], params[2]
    lamb = torch.exp(torch.complex(-nu, theta))
    xr, xi = gamma * u[0], gamma * u[1]
    fr, fi = lamb.real.expand_as(xr), lamb.imag.expand_as(xi)
    if side.rnn_start is not None:
        fr, fi = fr * (1 - side.rnn_start), fi * (1 - side.rnn_start)
    hr, hi = ops.lru_scan(xr, xi, fr, fi, None, None, side.grad_detach)
    out = ops.ensemble_linear(torch.stack((hr, hi), 0), p[pre + 'middle_proj.weight'], p[pre + 'middle_proj.bias'], desire_ndim=4)
    return ff_block(p, pre + 'ff.', out[0] - out[1] + u[2])


def mamba_mixer(p, pre, x, side: Side, cfg):
    """Mamba.forward_sequential, unfused branch (the one d_conv > 4 takes).
    ref: offpolicy_rnn/models/smamba/mamba.py:166-255"""
    Bsz, L, _ = x.shape
    Wx = p[pre + 'x_proj.weight']
    N = cfg['d_state']
    R = Wx.shape[0] - 2 * N
    xz = F.linear(x, p[pre + 'in_proj.weight']).transpose(1, 2)            # [B, 2D, L]     :175-179
    Dm = xz.shape[1] // 2
    xs, z = xz[:, :Dm], xz[:, Dm:]
    mask = None if side.mask is None else side.mask.transpose(-2, -1)      # :180-181
    xs = ops.causal_conv1d_silu(xs, p[pre + 'conv1d.weight'], p[pre + 'conv1d.bias'], mask)   # :210-212
    x_dbl = F.linear(xs.transpose(1, 2).reshape(Bsz * L, Dm), Wx)          # :231
    dt, Bm, Cm = torch.split(x_dbl, [R, N, N], dim=-1)
    dt = (p[pre + 'dt_proj.weight'] @ dt.t()).reshape(Dm, Bsz, L).transpose(0, 1)   # :233-234
    Bm = Bm.reshape(Bsz, L, N).transpose(1, 2)
    Cm = Cm.reshape(Bsz, L, N).transpose(1, 2)
    start = None if side.rnn_start is None else side.rnn_start.transpose(-2, -1).expand(Bsz, Dm, L)  # :182-183
    A = -torch.exp(p[pre + 'A_log'].float())                               # :187
    y = ops.selective_scan(xs, dt, A, Bm, Cm, start, p[pre + 'D'].float(), z=z,
                           delta_bias=p[pre + 'dt_proj.bias'].float(), delta_softplus=True)   # :238-250
    return F.linear(y.transpose(1, 2), p[pre + 'out_proj.weight'])         # :252


def smamba_layer(p, pre, x, side: Side, layer_id: str):
    """BlockList.forward with fused_add_norm semantics, norm eps 1e-8.
    ref: offpolicy_rnn/models/smamba/mamba.py:382-412 (Block), :492-526 (BlockList), eps :425"""
    cfg = parse_smamba(layer_id)
    eps = 1e-8
    residual = None
    for i in range(cfg['blocks']):
        b = f'{pre}layers.{i}.'
        hs, residual = ops.add_norm(x, p[b + 'norm.weight'], p.get(b + 'norm.bias'), residual, eps, prenorm=True,
                                    is_rms=cfg['rms'])
        x = mamba_mixer(p, b + 'mixer.', hs, side, cfg)
    if not cfg['ff']:
        x = ops.add_norm(x, p[pre + 'norm_f.weight'], p.get(pre + 'norm_f.bias'), residual, eps, prenorm=False,
                         is_rms=cfg['rms'])
        return F.linear(x, p[pre + 'head.weight'])
    x = x + residual
    return ff_block(p, pre + 'head.', x, eps)


def smamba_step(p, pre, x, hidden, layer_id: str):
    """Rollout path: ONE time step through every block's Mamba.step with the carried hidden
    [1, B, blocks * (D*d_conv + D*d_state)] (per block: conv window first, then SSM state).  rnn_start / mask are not
    consulted on this path.  ref: offpolicy_rnn/models/smamba/mamba.py:133-159 (dispatch), :257-305 (step),
    :492-526 (BlockList: hidden chunked per block, outputs concatenated)."""
    cfg = parse_smamba(layer_id)
    N, Kc, eps = cfg['d_state'], cfg['d_conv'], 1e-8
    Bsz = x.shape[0]
    hs = torch.chunk(hidden, cfg['blocks'], dim=-1)
    residual, outs = None, []
    for i in range(cfg['blocks']):
        b = f'{pre}layers.{i}.'
        m = b + 'mixer.'
        xin, residual = ops.add_norm(x, p[b + 'norm.weight'], p.get(b + 'norm.bias'), residual, eps, prenorm=True, is_rms=cfg['rms'])
        xz = F.linear(xin.squeeze(1), p[m + 'in_proj.weight'])
        Dm = xz.shape[-1] // 2
        xs, z = xz[:, :Dm], xz[:, Dm:]
        conv = hs[i][0, :, :Dm * Kc].reshape(Bsz, Dm, Kc)
        ssm = hs[i][0, :, Dm * Kc:].reshape(Bsz, Dm, N)
        conv = torch.cat((conv[:, :, 1:], xs.unsqueeze(-1)), dim=-1)                       # :264-266
        xs = F.silu((conv * p[m + 'conv1d.weight'][:, 0, :]).sum(-1) + p[m + 'conv1d.bias'])
        x_db = F.linear(xs, p[m + 'x_proj.weight'])
        R = x_db.shape[-1] - 2 * N
        dt, Bm, Cm = torch.split(x_db, [R, N, N], dim=-1)
        dt = F.softplus(F.linear(dt, p[m + 'dt_proj.weight']) + p[m + 'dt_proj.bias'])      # :283,289
        A = -torch.exp(p[m + 'A_log'].float())
        ssm = ssm * torch.exp(dt[..., None] * A) + xs[..., None] * (dt[..., None] * Bm[:, None, :])   # :290-292
        y = (ssm * Cm[:, None, :]).sum(-1) + p[m + 'D'] * xs
        x = F.linear(y * F.silu(z), p[m + 'out_proj.weight']).unsqueeze(1)
        outs.append(torch.cat((conv.reshape(1, Bsz, -1), ssm.reshape(1, Bsz, -1)), dim=-1))
    if not cfg['ff']:
        x = ops.add_norm(x, p[pre + 'norm_f.weight'], p.get(pre + 'norm_f.bias'), residual, eps, prenorm=False, is_rms=cfg['rms'])
        return F.linear(x, p[pre + 'head.weight']), torch.cat(outs, dim=-1)
    return ff_block(p, pre + 'head.', x + residual, eps), torch.cat(outs, dim=-1)


def gilr_lstm_layer(p, pre, x, side: Side, hidden=None):
    """ref: offpolicy_rnn/models/gilr_lstm/gilr_lstm.py:39-75 (CPU branch: scan_cpu with the carried halves)"""
    u = ops.ensemble_linear(x, p[pre + 'in_proj.weight'], p[pre + 'in_proj.bias'], desire_ndim=4)
    C = u.shape[-1]
    if hidden is None:
        h_pre = h_mid = None
    else:
        h_pre, h_mid = torch.chunk(hidden.transpose(0, 1), 2, -1)
    f, v = torch.sigmoid(u[1]), torch.tanh(u[0])
    if side.rnn_start is not None:
        f = f * (1 - side.rnn_start)
    v, h_pre = ops.gilr_scan(v, f, h_pre)
    g = ops.ensemble_linear(v, p[pre + 'middle_proj.weight'], p[pre + 'middle_proj.bias'], desire_ndim=4)
    f, i, o, z = torch.sigmoid(g[0]), torch.sigmoid(g[1]), torch.sigmoid(g[2]), torch.tanh(g[3])
    if side.rnn_start is not None:
        f = f * (1 - side.rnn_start)
    out, h_mid = ops.gilr_scan(i * z, f, h_mid)
    out = F.linear(out * o, p[pre + 'out_proj.weight'], p[pre + 'out_proj.bias'])
    return out, torch.cat((h_pre, h_mid), dim=-1).transpose(0, 1)


def conv1d_layer(p, pre, x, side: Side, layer_id: str, hidden=None):
    """ref: offpolicy_rnn/models/conv1d/conv1d.py:26-52 (explicit left state, padding 0, no activation, FF tail)"""
    Kc = int(layer_id.split('_')[-1]) if '_' in layer_id else 4
    Bsz, L, C = x.shape
    h = torch.zeros((Bsz, Kc - 1, C)) if hidden is None else hidden.reshape(Bsz, Kc - 1, C)
    xm = x if side.mask is None else x * side.mask
    x_in = torch.cat((h, xm), dim=-2)
    y = F.conv1d(x_in.transpose(-2, -1), p[pre + 'conv1d.weight'], p[pre + 'conv1d.bias'], groups=C)[:, :, :L].transpose(-2, -1)
    h_out = x_in[:, -(Kc - 1):, :].reshape(Bsz, 1, -1)
    return ff_block(p, pre + 'ff.', y), h_out


def parse_s6(layer_id: str):
    """ref: rnn_base.py:118-136"""
    cfg = dict(d_state=16, d_conv=4, ff=True)
    for tok in layer_id.split('_')[1:]:
        if tok.startswith('s'):
            cfg['d_state'] = int(tok[1:])
        elif tok.startswith('c'):
            cfg['d_conv'] = int(tok[1:])
        elif tok.startswith('no') and tok[2:] == 'ff':
            cfg['ff'] = False
    return cfg


def s6_layer(p, pre, x, side: Side, layer_id: str, hidden=None):
    """MambaResidualBlock: RMSNorm -> MambaBlock -> + x -> feed-forward tail; returns (out, hidden [B,1,D*N+(K-1)*D]).
    ref: offpolicy_rnn/models/s6/mamba.py:41-67 (block), :146-191 (mixer), :133-144 (conv with explicit left
    state, padding 0), :193-237 (ssm), RMSNorm :240-251 (eps 1e-5), PositionWiseFeedForward :256-267."""
    cfg = parse_s6(layer_id)
    N, K = cfg['d_state'], cfg['d_conv']
    Bsz, L, _ = x.shape
    m = pre + 'mixer.'
    rms = lambda t, w: t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + 1e-5) * w
    xz = F.linear(rms(x, p[pre + 'norm.weight']), p[m + 'in_proj.weight'])
    Dm = xz.shape[-1] // 2
    xs, res = xz[..., :Dm], xz[..., Dm:]
    if hidden is None:
        h_ssm, h_conv = torch.zeros((Bsz, Dm, N)), torch.zeros((Bsz, K - 1, Dm))
    else:
        h_ssm, h_conv = torch.split(hidden, [Dm * N, Dm * (K - 1)], dim=-1)
        h_ssm, h_conv = h_ssm.reshape(Bsz, Dm, N), h_conv.reshape(Bsz, K - 1, Dm)
    if side.mask is not None:
        xs = xs * side.mask
    x_in = torch.cat((h_conv, xs), dim=-2)                                        # :139
    xs = F.conv1d(x_in.transpose(1, 2), p[m + 'conv1d.weight'], p[m + 'conv1d.bias'], groups=Dm)[:, :, :L].transpose(1, 2)
    h_conv = x_in[:, -(K - 1):, :]
    xs = F.silu(xs)
    x_dbl = F.linear(xs, p[m + 'x_proj.weight'])
    R = x_dbl.shape[-1] - 2 * N
    dt, Bm, Cm = torch.split(x_dbl, [R, N, N], dim=-1)
    delta = F.softplus(F.linear(dt, p[m + 'dt_proj.weight'], p[m + 'dt_proj.bias']))
    start = side.rnn_start if side.rnn_start is not None else torch.zeros((Bsz, L, 1))
    y, h_ssm = ops.s6_scan(xs, delta, -torch.exp(p[m + 'A_log'].float()), Bm, Cm, p[m + 'D'].float(), start, h_ssm)
    out = F.linear(y * F.silu(res), p[m + 'out_proj.weight']) + x
    h_out = torch.cat((h_ssm.reshape(Bsz, 1, -1), h_conv.reshape(Bsz, 1, -1)), dim=-1)
    if cfg['ff']:
        return ff_block(p, pre + 'ff.', out), h_out
    return F.linear(rms(out, p[pre + 'norm_f.weight']), p[pre + 'ff.weight']), h_out


def gru_layer(p, pre, x, side: Side, hidden=None):
    """torch.nn.GRU(batch_first=True); initial state zero unless one is carried in ([1, B, H]).  ref: rnn_base.py:59,245-247,454"""
    width = p[pre + 'weight_hh_l0'].shape[1]
    h0 = torch.zeros((1, x.shape[0], width), dtype=x.dtype) if hidden is None else hidden
    flat = [p[pre + 'weight_ih_l0'], p[pre + 'weight_hh_l0'], p[pre + 'bias_ih_l0'], p[pre + 'bias_hh_l0']]
    out, _ = torch._VF.gru(x, h0, flat, True, 1, 0.0, False, False, True)
    return out


def cgpt_layer(p, pre, x, side: Side, layer_id: str):
    """TransformerDecoder on the rows' packed sequences (eval / p = 0 semantics).
    ref: offpolicy_rnn/models/rnn_base.py:222-236 (grammar), :437-452 (call), flash_attention/TransformerFlashAttention.py:104-121"""
    from . import attention as OA
    nhead, ln = 8, True
    for tok in layer_id.split('_')[1:]:
        if tok.startswith('h'):
            nhead = int(tok[1:])
        elif tok.startswith('rms'):
            ln = False
    seq = side.attention_concat_mask
    seq = None if seq is None else seq.detach().cpu().numpy().astype('int64')
    sub = {k[len(pre):]: v for k, v in p.items() if k.startswith(pre)}
    return OA.decoder_forward(sub, x, seq, nhead, ln)


def rnn_base(p: Dict[str, torch.Tensor], layer_types: List[str], acts: List[str], x, side: Optional[Side] = None,
             desire_ndim=None):
    """RNNBase.meta_forward.  ref: offpolicy_rnn/models/rnn_base.py:397-472"""
    side = side or Side()
    for i, (lt, act) in enumerate(zip(layer_types, acts)):
        pre = f'layer_list.{i}.'
        if lt == 'fc':
            x = F.linear(x, p[pre + 'weight'], p[pre + 'bias'])
        elif lt.startswith('efc'):
            x = ops.ensemble_linear(x, p[pre + 'weight'], p[pre + 'bias'], desire_ndim)
        elif lt == 'gilr':
            x = gilr_layer(p, pre, x, side)
        elif lt == 'gilr_lstm':
            x, h = gilr_lstm_layer(p, pre, x, side, side.h0.get(i))
            side.h_out = h
        elif lt.startswith('conv1d'):
            x, h = conv1d_layer(p, pre, x, side, lt, side.h0.get(i))
            side.h_out = h
        elif lt == 'lru':
            x = lru_layer(p, pre, x, side)
        elif lt.startswith('smamba'):
            x = smamba_layer(p, pre, x, side, lt)
        elif lt.startswith('mamba'):
            x, h = s6_layer(p, pre, x, side, lt, side.h0.get(i))
            side.h_out = h
        elif lt == 'gru':
            x = gru_layer(p, pre, x, side, side.h0.get(i))
        elif lt.startswith('cgpt'):
            x = cgpt_layer(p, pre, x, side, lt)
        else:
            raise NotImplementedError(lt)
        if '+' in act:
            norm, a = act.split('+')
            x = F.layer_norm(x, x.shape[-1:], p[f'activation_list.{i}.0.weight'], p[f'activation_list.{i}.0.bias'])
            x = ACT[a](x)
        else:
            x = ACT[act](x)
    return x


class ModelSpec:
    """The constructor kwargs the reference's make_*_model takes (ref: algorithm/sac.py:199-239)."""

    def __init__(self, state_dim, action_dim, embedding_size, embedding_hidden, embedding_activations,
                 embedding_layer_type, uni_model_hidden, uni_model_activations, uni_model_layer_type,
                 last_state_input=True, last_action_input=True, reward_input=False, **_):
        self.state_dim, self.action_dim = state_dim, action_dim
        self.emb_types, self.emb_acts = list(embedding_layer_type), list(embedding_activations)
        self.uni_types = list(uni_model_layer_type)
        self.uni_acts = list(uni_model_activations)
        self.last_state_input, self.last_action_input, self.reward_input = last_state_input, last_action_input, reward_input
        self.map_act = embedding_activations[-1]


def embedding_input(sd, spec: ModelSpec, state, lst_state, lst_action, reward):
    """ref: contextual_sac_policy_single_head.py:81-90 / contextual_sac_value.py:90-99"""
    lin = lambda name, x: F.linear(x, sd[name]['weight'], sd[name]['bias'])
    parts = [lin('state_encoder', state)]
    if spec.last_state_input:
        parts.append(lin('last_obs_encoder', lst_state))
    if spec.last_action_input:
        parts.append(lin('last_act_encoder', lst_action))
    if spec.reward_input:
        parts.append(lin('reward_encoder', reward))
    return torch.cat(parts, dim=-1)


def contextual_forward(sd, spec: ModelSpec, emb_in, uni_in, side, detach_embedding, desire_ndim=None):
    """ContextualModel.meta_forward.  ref: contextual_model.py:57-116"""
    emb = rnn_base(sd['embedding_model'], spec.emb_types, spec.emb_acts, emb_in, side)
    if detach_embedding:
        emb = emb.detach()
    if 'uni_input_mapping_network' in sd:
        m = sd['uni_input_mapping_network']
        uni_in = ACT[spec.map_act](F.linear(uni_in, m['layer_list.0.weight'], m['layer_list.0.bias']))
    out = rnn_base(sd['universal_model'], spec.uni_types, spec.uni_acts, torch.cat((uni_in, emb), dim=-1), side,
                   desire_ndim)
    return out, emb


def policy_forward(sd, spec: ModelSpec, state, lst_state, lst_action, side, reward=None, noise=None, td3=False,
                   sample_std=0.1):
    """ref: contextual_sac_policy_single_head.py:92-123; TD3: contextual_td3_policy.py:18-36.
    `noise` replaces torch.randn_like so that runs are comparable."""
    out, emb = contextual_forward(sd, spec, embedding_input(sd, spec, state, lst_state, lst_action, reward), state, side, False)
    if td3:
        mean = torch.tanh(out)
        sample = torch.clamp(mean + noise * sample_std, -1, 1)
        return mean, emb, sample, torch.zeros_like(sample)
    logstd, mean_raw = out.chunk(2, dim=-1)
    mean, sample, logp = ops.tanh_gaussian(mean_raw, logstd, noise)
    return mean, emb, sample, logp


def value_forward(sd, spec: ModelSpec, state, lst_state, lst_action, action, side, reward=None, detach_embedding=False):
    """ref: contextual_sac_value.py:101-119 (state_action + meta_forward), desire_ndim = 4
    (ref: sac_full_length_rnn_ensembleQ.py:25-32)"""
    lin = lambda name, x: F.linear(x, sd[name]['weight'], sd[name]['bias'])
    sa = torch.cat((lin('state_input_encoder_q', state), lin('action_input_encoder_q', action)), dim=-1)
    sa = ACT[spec.map_act](sa)
    q, emb = contextual_forward(sd, spec, embedding_input(sd, spec, state, lst_state, lst_action, reward), sa, side,
                                detach_embedding, desire_ndim=4)
    return q, emb


def policy_forward_discrete(sd, spec: ModelSpec, state, lst_state, lst_action, side, reward=None):
    """Categorical policy: softmax of the head output mixed with a 0.01 floor, renormalised; returns (mode, emb, log_probs).
    ref: contextual_sac_discrete_policy.py:88-118"""
    out, emb = contextual_forward(sd, spec, embedding_input(sd, spec, state, lst_state, lst_action, reward), state, side, False)
    probs = (out - out.max(dim=-1, keepdim=True).values).exp()
    probs = probs / probs.sum(dim=-1, keepdim=True)
    probs = probs + 0.01
    probs = probs / probs.sum(dim=-1, keepdim=True)
    return probs.argmax(dim=-1, keepdim=True), emb, torch.log(probs)


def value_forward_discrete(sd, spec: ModelSpec, state, lst_state, lst_action, side, reward=None, detach_embedding=False):
    """One Q per action from act(Linear(state)) and the embedding; the action argument of the reference's forward is
    ignored there.  ref: contextual_sac_discrete_value.py:98-126"""
    lin = lambda name, x: F.linear(x, sd[name]['weight'], sd[name]['bias'])
    se = ACT[spec.map_act](lin('state_input_encoder_q', state))
    return contextual_forward(sd, spec, embedding_input(sd, spec, state, lst_state, lst_action, reward), se, side,
                              detach_embedding, desire_ndim=4)
