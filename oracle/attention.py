"""ORACLE (test infrastructure, not product code) -- fp32 CPU restatement of the cgpt encoder layer.

ref: offpolicy_rnn/models/flash_attention/TransformerFlashAttention.py:29-121 (RMSNorm, PositionWiseFeedForward,
     DecoderLayer, TransformerDecoder); the attention itself is third-party in the reference (flash-attn 2.x,
     `flash_attn_varlen_qkvpacked_func(..., causal=True, alibi_slopes=...)` behind flash_attn.modules.mha.MHA,
     requirement.txt:7, unpinned; 2.8.3 in this image).  Its published semantics are restated here:
     score_ij = softmax_scale * q_i . k_j - slope_h * (i - j) for j <= i inside one sequence, softmax over j,
     slopes from get_alibi_slopes (2^(-8 (h+1) / n) for power-of-two n).
PINNED: tests/golden/layer_cgpt_{ln,rms}.npz, step_cgpt.npz and update_sac_cgpt.npz are outputs of the UNMODIFIED
reference (staged copy oracle/_ref) run with flash-attn 2.8.3 on a B200 by tests/golden/make_golden_gpu.py;
tests/test_oracle_golden.py checks this restatement against them at the bf16 tolerance (1e-2), and
tests/test_attn_gpu.py checks the CUDA path against the same files and against flash-attn's kernel directly.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def alibi_slopes(nheads: int):
    def pow2(n):
        start = 2 ** (-(2 ** -(math.log2(n) - 3)))
        return [start * start ** i for i in range(n)]
    if math.log2(nheads).is_integer():
        return pow2(nheads)
    closest = 2 ** math.floor(math.log2(nheads))
    return pow2(closest) + alibi_slopes(2 * closest)[0::2][:nheads - closest]


def attention_varlen(qkv: torch.Tensor, starts, lens, slopes, softmax_scale: float) -> torch.Tensor:
    """qkv [T, 3, H, hd] -> out [T, H*hd]; tokens outside every sequence give zeros (what pad_input leaves)."""
    T, _, H, hd = qkv.shape
    out = torch.zeros((T, H * hd), dtype=qkv.dtype)
    sl = torch.as_tensor(slopes, dtype=qkv.dtype).view(H, 1, 1)
    pieces = {}
    for s, n in zip(starts, lens):
        q, k, v = (qkv[s:s + n, i].transpose(0, 1) for i in range(3))           # [H, n, hd]
        pos = torch.arange(n)
        rel = (pos.view(n, 1) - pos.view(1, n)).to(qkv.dtype)                    # i - j
        sc = softmax_scale * q @ k.transpose(1, 2) - sl * rel
        sc = sc.masked_fill(rel.unsqueeze(0) < 0, float('-inf'))
        o = torch.softmax(sc, dim=-1) @ v                                        # [H, n, hd]
        pieces[s] = o.transpose(0, 1).reshape(n, H * hd)
    if pieces:
        idx = torch.cat([torch.arange(s, s + p.shape[0]) for s, p in pieces.items()])
        out = out.index_put((idx,), torch.cat(list(pieces.values())))
    return out


def rms_norm(x, w, eps=1e-5):
    """ref: TransformerFlashAttention.py:29-40"""
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w


def decoder_forward(sd, x, seqlens, nhead: int, ln: bool, prefix: str = ''):
    """TransformerDecoder.forward in eval mode / p = 0 (ref :104-121) on x [B, L, C] with row-packed sequence
    lengths `seqlens` [B, <=L] (host ints; None = one full sequence per row)."""
    B, L, C = x.shape
    hd = C // nhead
    starts, lens, rows = [], [], []
    for b in range(B):
        pos = 0
        for n in ([L] if seqlens is None else list(seqlens[b])):
            n = int(n)
            if n > 0:
                starts.append(b * L + pos); lens.append(n); pos += n
        rows.append(pos)
    norm = (lambda t, p: F.layer_norm(t, (C,), sd[p + '.weight'], sd[p + '.bias'], 1e-5)) if ln else \
           (lambda t, p: rms_norm(t, sd[p + '.weight']))
    h = x.reshape(B * L, C)
    n_layer = 1 + max(int(k[len(prefix):].split('.')[1]) for k in sd if k.startswith(prefix + 'decoder_layers.'))
    for i in range(n_layer):
        p = f'{prefix}decoder_layers.{i}.'
        a = norm(h, p + 'mha_norm')
        qkv = F.linear(a, sd[p + 'mha.Wqkv.weight'], sd[p + 'mha.Wqkv.bias']).view(B * L, 3, nhead, hd)
        o = attention_varlen(qkv, starts, lens, alibi_slopes(nhead), 1.0 / math.sqrt(hd))
        h = F.linear(o, sd[p + 'mha.out_proj.weight'], sd[p + 'mha.out_proj.bias']) + h
        f = norm(h, p + 'ffn_norm')
        f = F.linear(F.gelu(F.linear(f, sd[p + 'ffn.fc1.weight'], sd[p + 'ffn.fc1.bias'])), sd[p + 'ffn.fc2.weight'], sd[p + 'ffn.fc2.bias'])
        h = f + h
    h = F.linear(norm(h, prefix + 'output_ln'), sd[prefix + 'output_fc.weight'], sd[prefix + 'output_fc.bias']).view(B, L, C)
    keep = torch.zeros((B, L, 1), dtype=x.dtype)
    for b, n in enumerate(rows):
        keep[b, :n] = 1
    return h * keep
