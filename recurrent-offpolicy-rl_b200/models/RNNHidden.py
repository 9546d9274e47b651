"""Per-layer recurrent state list + the side-band the update path threads through every encoder call
(`rnn_start`, `mask`, `attention_concat_mask`, `grad_detach`).

API-compatible with the reference's RNNHidden for everything the update path and the model classes
touch (ref: offpolicy_rnn/models/RNNHidden.py:12-62 constructor + side-band, :93-130 init / slicing,
:376-385 `+`).  States are `[1, B, dim]` tensors (batch_first=False) or `[B, L, dim]` full outputs
(batch_first=True); `lstm` states are (h, c) tuples.
"""
from __future__ import annotations

import copy
from typing import List, Sequence, Tuple, Union

import torch

State = Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor], None, object]

_PLAIN = ('gru', 'lru', 'gilr', 'cgru', 'gilr_lstm', 'mamba', 'conv1d', 'smamba', 'transformer')


def _is_recurrent(name: str) -> bool:
    if name in _PLAIN:
        return True
    if name.startswith('e') and name[1:].split('-')[0] in _PLAIN:
        return True
    return name.startswith(('conv1d', 'econv1d', 'mamba', 'smamba', 'transformer'))


class AttentionCache:
    """Stand-in for the reference's InferenceParams (kv-cache bookkeeping of cgpt/gpt layers,
    ref: offpolicy_rnn/models/flash_attention/TransformerFlashAttention.py:12-27)."""

    def __init__(self, max_seqlen: int, max_batch_size: int):
        self.max_seqlen, self.max_batch_size = max_seqlen, max_batch_size
        self.seqlen_offset = 0
        self.batch_size_offset = 0
        self.key_value_memory_dict = {}
        self.lengths_per_sample = None

    def reset(self, max_seqlen, max_batch_size):
        self.max_seqlen, self.max_batch_size, self.seqlen_offset = max_seqlen, max_batch_size, 0


InferenceParams = AttentionCache


class RNNHidden:
    SUPPORTED_RNN_TYPES = list(_PLAIN)

    def __init__(self, rnn_num: int, rnn_types: Sequence[str], device=torch.device('cpu'), batch_first: bool = False):
        assert len(rnn_types) == rnn_num, 'number of rnn layers should be equal to the rnn types'
        self._rnn_num = rnn_num
        self._rnn_types: List[str] = list(rnn_types)
        self._device = torch.device(device) if not isinstance(device, torch.device) else device
        self._batch_first = batch_first
        self._data: List[State] = []
        self._rnn_start = self._mask = self._attention_concat_mask = self._grad_detach = None

    # ---- side-band -------------------------------------------------------------------------------
    def set_rnn_start(self, v): self._rnn_start = v
    def set_mask(self, v): self._mask = v
    def set_attention_concat_mask(self, v): self._attention_concat_mask = v
    def set_grad_detach(self, v): self._grad_detach = v
    rnn_start = property(lambda self: self._rnn_start)
    mask = property(lambda self: self._mask)
    attention_concat_mask = property(lambda self: self._attention_concat_mask)
    grad_detach = property(lambda self: self._grad_detach)
    size = property(lambda self: len(self._data))
    device = property(lambda self: self._device)
    capacity = property(lambda self: self._rnn_num)

    def _carry_sideband(self, other: "RNNHidden") -> "RNNHidden":
        other._rnn_start, other._mask = self._rnn_start, self._mask
        other._attention_concat_mask, other._grad_detach = self._attention_concat_mask, self._grad_detach
        return other

    # ---- container -------------------------------------------------------------------------------
    def append(self, state: State, rnn_type=None) -> None:
        assert len(self._data) < self._rnn_num, 'hidden num exceeds the number of RNN layers'
        if rnn_type is not None:
            assert rnn_type == self._rnn_types[len(self._data)]
        self._data.append(state)

    def __len__(self): return len(self._data)

    def __getitem__(self, key):
        if isinstance(key, slice):
            out = RNNHidden(len(self._data[key]), self._rnn_types[key], self._device, self._batch_first)
            out._data = self._data[key]
            return self._carry_sideband(out)
        return self._data[key]

    def __setitem__(self, key, value): self._data[key] = value

    def __add__(self, other):
        if other is None:
            return self
        if not isinstance(other, RNNHidden):
            return NotImplemented
        out = RNNHidden(self._rnn_num + other._rnn_num, self._rnn_types + other._rnn_types, self._device, self._batch_first)
        out._data = self._data + other._data
        return out

    def __copy__(self):
        out = RNNHidden(self._rnn_num, self._rnn_types, self._device, self._batch_first)
        out._data = self._data
        return self._carry_sideband(out)

    def __deepcopy__(self, memo):
        out = RNNHidden(self._rnn_num, self._rnn_types, self._device, self._batch_first)
        for s in self._data:
            if isinstance(s, tuple):
                out._data.append(tuple(t.clone() for t in s))
            elif isinstance(s, torch.Tensor):
                out._data.append(s.clone())
            else:
                out._data.append(copy.deepcopy(s))
        for n in ('_rnn_start', '_mask', '_attention_concat_mask', '_grad_detach'):
            v = getattr(self, n)
            setattr(out, n, None if v is None else v.clone())
        return out

    # ---- construction helpers ----------------------------------------------------------------------
    @torch.no_grad()
    def init_hidden_by_type(self, rnn_type: str, batch_size: int, unit_num: int, device):
        if rnn_type == 'lstm':
            return (torch.zeros((1, batch_size, unit_num), device=device), torch.zeros((1, batch_size, unit_num), device=device))
        if rnn_type.startswith(('gpt', 'cgpt')):
            return AttentionCache(max_seqlen=unit_num, max_batch_size=batch_size)
        if _is_recurrent(rnn_type):
            z = torch.zeros((1, batch_size, unit_num), device=device)
            z._rorl_zero = True      # lets layers skip the carried-state term without a host sync
            return z
        raise NotImplementedError(f'rnn type: {rnn_type} has not been implemented!!')

    @torch.no_grad()
    def init_random_hidden_by_type(self, rnn_type: str, batch_size: int, unit_num: int, device):
        u = lambda: torch.rand((1, batch_size, unit_num), device=device) * 2 - 1
        if rnn_type == 'lstm':
            return (u(), u())
        if rnn_type.startswith(('gpt', 'cgpt')):
            return AttentionCache(max_seqlen=unit_num, max_batch_size=batch_size)
        if _is_recurrent(rnn_type):
            return u()
        raise NotImplementedError(f'rnn type: {rnn_type} has not been implemented!!')

    # ---- transformations used by the rollout / slicing code of the reference ------------------------
    def _map(self, fn):
        for i, s in enumerate(self._data):
            if isinstance(s, tuple):
                self._data[i] = tuple(fn(t) for t in s)
            elif isinstance(s, torch.Tensor):
                self._data[i] = fn(s)
            elif s is not None:
                raise NotImplementedError('attention caches do not support this operation')

    def to_device(self, device):
        if self._device != device:
            self._device = device
            for i, s in enumerate(self._data):
                if isinstance(s, tuple):
                    self._data[i] = tuple(t.to(device) for t in s)
                elif isinstance(s, torch.Tensor):
                    self._data[i] = s.to(device)

    def hidden_detach_(self): self._map(lambda t: t.detach())
    def hidden_state_slice_(self, start, end): self._map(lambda t: t[:, start:end])
    def hidden_state_sample_(self, idxes): self._map(lambda t: t[:, idxes])
    def hidden_state_mask_(self, masks): self._map(lambda t: t.squeeze(0)[masks].unsqueeze(0))

    def _copied(self, name, *a):
        out = copy.deepcopy(self)
        getattr(out, name)(*a)
        return out

    def hidden_detach(self): return self._copied('hidden_detach_')
    def hidden_state_slice(self, start, end): return self._copied('hidden_state_slice_', start, end)
    def hidden_state_sample(self, idxes): return self._copied('hidden_state_sample_', idxes)
    def hidden_state_mask(self, masks): return self._copied('hidden_state_mask_', masks)

    @property
    def hidden_batch_size(self) -> int:
        if not self._data or self._data[0] is None:
            return 0
        s = self._data[0]
        return (s[0] if isinstance(s, tuple) else s).shape[1]

    def elementwise_append(self, other: "RNNHidden") -> None:
        assert other._rnn_num == self._rnn_num and other.size == self.size and not self._batch_first
        for i, (a, b) in enumerate(zip(self._data, other._data)):
            if a is None:
                self._data[i] = b
            elif isinstance(a, tuple):
                self._data[i] = tuple(torch.cat((x, y), dim=1) for x, y in zip(a, b))
            else:
                self._data[i] = torch.cat((a, b), dim=1)

    def elementwise_pop(self, pop_num: int = 1) -> None:
        for i, a in enumerate(self._data):
            if a is None:
                continue
            first = a[0] if isinstance(a, tuple) else a
            if first.shape[1] <= pop_num:
                self._data[i] = None
            elif isinstance(a, tuple):
                self._data[i] = tuple(x[:, pop_num:, :] for x in a)
            else:
                self._data[i] = a[:, pop_num:, :]

    def full_rnn_insert_init_hidden_(self, init_hidden=None, pop_final_hidden=False):
        for i, full in enumerate(self._data):
            h0 = (torch.zeros((full.shape[0], 1, full.shape[-1]), device=full.device) if init_hidden is None
                  else init_hidden[i].squeeze(0).unsqueeze(1))
            self._data[i] = torch.cat((h0, full[..., :-1, :] if pop_final_hidden else full), dim=1)
        return self

    def reshape_full_rnn_output_to_hidden_(self, target_traj_len: int):
        assert self._batch_first
        for i, full in enumerate(self._data):
            idx = [k * target_traj_len for k in range(full.shape[1] // target_traj_len)]
            picked = full[:, idx, :].transpose(0, 1)
            self._data[i] = picked.reshape((1, picked.shape[0] * picked.shape[1], picked.shape[2]))
        self._batch_first = False
        return self

    def reshape_full_rnn_output_to_hidden(self, target_traj_len: int):
        return copy.deepcopy(self).reshape_full_rnn_output_to_hidden_(target_traj_len)

    def __str__(self):
        return '\n'.join(f'RNN hidden {i + 1}/{len(self._data)}: {getattr(s, "shape", s)}' for i, s in enumerate(self._data))
