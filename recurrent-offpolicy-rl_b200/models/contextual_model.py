"""Context encoder ("embedding network") + universal network pair shared by the policy and value models.

API-compatible with the reference's ContextualModel (ref: offpolicy_rnn/models/contextual_model.py):
a plain Python object (not an nn.Module) holding `embedding_network`, `uni_network`, the optional
`uni_input_mapping_network`, and the name -> module registry `contextual_modules` that the RESeL
optimizer split, save/load, soft update and state_dict all iterate (ref :9-44, :135-173).
`fix_rnn_length` > 0 (sliced-sequence algorithms) is outside the full-length update path and rejected.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch

from .RNNHidden import RNNHidden
from .mlp_base import MLPBase
from .rnn_base import RNNBase


class ContextualModel:
    def __init__(self, embedding_input_size: int, embedding_size: int, embedding_hidden: List[int],
                 embedding_activations: List[str], embedding_layer_type: List[str], uni_model_input_size: int,
                 uni_model_output_size: int, uni_model_hidden: List[int], uni_model_activations: List[str],
                 uni_model_layer_type: List[str], fix_rnn_length: int, name: str, uni_model_input_mapping_dim: int = 0,
                 uni_model_input_mapping_activation: str = 'linear'):
        if fix_rnn_length and fix_rnn_length > 0:
            raise NotImplementedError('fix_rnn_length > 0 belongs to the sliced-sequence algorithms, not the full-length update path')
        self.name = name
        self._fix_rnn_length = self.fix_rnn_length = 0
        self.embedding_size = embedding_size
        self.uni_model_input_mapping_dim = uni_model_input_mapping_dim
        self.embedding_network = RNNBase(embedding_input_size, embedding_size, embedding_hidden, embedding_activations,
                                         embedding_layer_type)
        uni_in = uni_model_input_size if uni_model_input_mapping_dim == 0 else uni_model_input_mapping_dim
        self.uni_network = RNNBase(embedding_size + uni_in, uni_model_output_size, uni_model_hidden,
                                   uni_model_activations, uni_model_layer_type)
        self.contextual_modules: Dict[str, torch.nn.Module] = {}
        self.contextual_register_rnn_base_module(self.embedding_network, 'embedding_model')
        self.contextual_register_rnn_base_module(self.uni_network, 'universal_model')
        if uni_model_input_mapping_dim > 0:
            self.uni_input_mapping_network = MLPBase(uni_model_input_size, uni_model_input_mapping_dim, [],
                                                     [uni_model_input_mapping_activation])
            self.contextual_register_rnn_base_module(self.uni_input_mapping_network, 'uni_input_mapping_network')
        else:
            self.uni_input_mapping_network = torch.nn.Identity()
        self.rnn_num = self.embedding_network.rnn_num + self.uni_network.rnn_num
        self.device = torch.device('cpu')
        self.dtype = torch.float32
        # forward() also returns every recurrent layer's output sequence (`full_rnn_memory`, ref: contextual_model.py:57-116).
        # A caller that never reads it (the update engine) clears this flag: the record is skipped and the activation that
        # follows a recurrent layer may then be fused into that layer's last kernel.
        self.need_full_hidden = True

    def contextual_register_rnn_base_module(self, module, module_name: str):
        self.contextual_modules[module_name] = module

    def parameters(self, recursive=True) -> List[torch.Tensor]:
        return [p for m in self.contextual_modules.values() for p in m.parameters(recursive)]

    def named_parameters(self):
        for k, m in self.contextual_modules.items():
            for n, p in m.named_parameters():
                yield f'{k}.{n}', p

    def rnn_parameters(self, recursive=True):
        return [p for m in self.contextual_modules.values() if hasattr(m, 'rnn_parameters') for p in m.rnn_parameters(recursive)]

    # ---- forward (ref: contextual_model.py:57-116) ---------------------------------------------------------
    def meta_forward(self, embedding_input, uni_model_input, rnn_memory=None, detach_embedding=False, embedded=None):
        """embedded: the result of `_meta_forward_embedding` computed earlier (e.g. on a side stream while another model's
        encoder runs); `embedding_input` is then ignored."""
        if rnn_memory is None:
            rnn_memory = self.make_init_state(1 if embedding_input.dim() == 2 else embedding_input.shape[0], embedding_input.device)
        emb, emb_mem, emb_full = embedded if embedded is not None else self._meta_forward_embedding(embedding_input, rnn_memory)
        if detach_embedding:
            emb = emb.detach()
        out, uni_mem, uni_full = self._meta_forward_uni_model(uni_model_input, emb, rnn_memory)
        return out, emb_mem + uni_mem, emb, (emb_full + uni_full if self.need_full_hidden else None)

    def _meta_forward_embedding(self, embedding_input, rnn_memory: Optional[RNNHidden]):
        mem = rnn_memory[:self.embedding_network.rnn_num] if rnn_memory is not None and len(rnn_memory) > 0 else None
        return self.embedding_network.meta_forward(embedding_input, mem, require_full_hidden=self.need_full_hidden)

    def _meta_forward_uni_model(self, uni_model_input, embedding, rnn_memory: Optional[RNNHidden]):
        uni_model_input = self.uni_input_mapping_network(uni_model_input)
        mem = rnn_memory[self.embedding_network.rnn_num:] if rnn_memory is not None and len(rnn_memory) > 0 else None
        if embedding.dim() - uni_model_input.dim() == 1:
            uni_model_input = uni_model_input.unsqueeze(0).repeat_interleave(repeats=embedding.shape[0], dim=0)
        return self.uni_network.meta_forward(torch.cat((uni_model_input, embedding), dim=-1), mem, require_full_hidden=self.need_full_hidden)

    def get_embedding(self, x, rnn_memory):
        return self._meta_forward_embedding(x, rnn_memory)

    # ---- housekeeping ------------------------------------------------------------------------------------------
    def to(self, device=None, dtype=None) -> None:
        if device is not None and self.device != device:
            self.device = device
            for m in self.contextual_modules.values():
                m.to(device)
        if dtype is not None and self.dtype != dtype:
            self.dtype = dtype
            for m in self.contextual_modules.values():
                m.to(dtype)

    def save(self, path: str, index: int = 0) -> None:
        for k, m in self.contextual_modules.items():
            full = os.path.join(path, f'{self.name}-{index}-{k}.pt')
            os.makedirs(os.path.dirname(full), exist_ok=True)
            torch.save(m.state_dict(), full)

    def load(self, path: str, index: int = 0, **kwargs) -> None:
        for k, m in self.contextual_modules.items():
            m.load_state_dict(torch.load(os.path.join(path, f'{self.name}-{index}-{k}.pt'), **kwargs))

    def copy_weight_from(self, src: "ContextualModel", tau: float) -> None:
        """target update: tau = 0 copies, tau = 1 keeps (ref: contextual_model.py:156-163)."""
        for k, m in self.contextual_modules.items():
            RNNBase._copy_weight_from(m, src.contextual_modules[k], tau)

    def state_dict(self, destination=None, prefix='', keep_vars=False):
        return {k: m.state_dict(destination=destination, prefix=prefix, keep_vars=keep_vars) for k, m in self.contextual_modules.items()}

    def load_state_dict(self, state_dict):
        for k, m in self.contextual_modules.items():
            m.load_state_dict(state_dict[k])

    def make_init_state(self, batch_size: int, device) -> RNNHidden:
        return self.embedding_network.make_init_state(batch_size, device) + self.uni_network.make_init_state(batch_size, device)

    def make_rnd_init_state(self, batch_size, device) -> RNNHidden:
        return self.embedding_network.make_rnd_init_state(batch_size, device) + self.uni_network.make_rnd_init_state(batch_size, device)

    def train(self, mode=True):
        for m in self.contextual_modules.values():
            m.train(mode)

    def eval(self):
        for m in self.contextual_modules.values():
            m.eval()

    def set_fix_length(self, enable: bool):
        if enable and self._fix_rnn_length:
            raise NotImplementedError

    def l2_norm_square(self) -> torch.Tensor:
        return sum(m.l2_norm_square() for m in self.contextual_modules.values() if hasattr(m, 'l2_norm_square'))
