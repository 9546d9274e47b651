"""Recurrent ensemble-Q value model.

API-compatible with ContextualSACValue / ContextualTD3Value
(ref: offpolicy_rnn/policy_value_models/contextual_sac_value.py:9-126, contextual_td3_value.py:9-17):
with `separate_encoder` the universal network sees ELU(cat(Linear(state), Linear(action))) and the two
extra encoders are registered as `state_input_encoder_q` / `action_input_encoder_q` (ref :48-57, :85-88).
forward returns `(Q [E, B, L, 1], embedding, rnn_memory, full_rnn_memory)`.
"""
from typing import Optional

import torch

from .. import kernels as K
from ..models.contextual_model import ContextualModel
from ..models.linear import Linear
from ..models.RNNHidden import RNNHidden
from ..models.rnn_base import ACTIVATIONS
from .contextual_sac_policy import _InputEncoders
from .utils import nearest_power_of_two, nearest_power_of_two_half


class ContextualSACValue(ContextualModel, _InputEncoders):
    def __init__(self, state_dim, action_dim, embedding_size, embedding_hidden, embedding_activations,
                 embedding_layer_type, uni_model_hidden, uni_model_activations, uni_model_layer_type, fix_rnn_length,
                 uni_model_input_mapping_dim: int = 0, reward_input=False, last_action_input=True, last_state_input=False,
                 separate_encoder=False, name='ContextualSACValue'):
        self.embedding_state_dim = state_dim
        if embedding_size == 'auto':
            embedding_size = nearest_power_of_two_half(state_dim)
        if uni_model_input_mapping_dim == 'auto':
            uni_model_input_mapping_dim = nearest_power_of_two(state_dim + action_dim)
        cum_dim = self._build_encoders(state_dim, action_dim, reward_input, last_action_input, last_state_input, separate_encoder)
        uni_in = state_dim + action_dim
        self.state_input_encoder = torch.nn.Identity()
        self.action_input_encoder = torch.nn.Identity()
        if uni_model_input_mapping_dim > 0 and separate_encoder:
            self.state_input_encoder = Linear(state_dim, uni_model_input_mapping_dim)
            self.action_input_encoder = Linear(action_dim, uni_model_input_mapping_dim)
            uni_in = uni_model_input_mapping_dim * 2
            uni_model_input_mapping_dim = 0
        ContextualModel.__init__(self, embedding_input_size=cum_dim, embedding_size=embedding_size,
                                 embedding_hidden=embedding_hidden, embedding_activations=embedding_activations,
                                 embedding_layer_type=embedding_layer_type, uni_model_input_size=uni_in,
                                 uni_model_output_size=1, uni_model_hidden=uni_model_hidden,
                                 uni_model_activations=uni_model_activations, uni_model_layer_type=uni_model_layer_type,
                                 fix_rnn_length=fix_rnn_length, uni_model_input_mapping_dim=uni_model_input_mapping_dim,
                                 uni_model_input_mapping_activation=embedding_activations[-1], name=name)
        self.uni_model_input_mapping_activation_func = ACTIVATIONS[embedding_activations[-1]]()
        self._register_encoders()
        if separate_encoder:
            self.contextual_register_rnn_base_module(self.state_input_encoder, 'state_input_encoder_q')
            self.contextual_register_rnn_base_module(self.action_input_encoder, 'action_input_encoder_q')
        self.state_dim, self.action_dim = state_dim, action_dim

    def state_action(self, state, action):
        if self.separate_encoder and isinstance(self.state_input_encoder, Linear):
            xs, mods = [state, action], [self.state_input_encoder, self.action_input_encoder]
            act = self.uni_model_input_mapping_activation_func
            if K.skinny_encoders_ok(xs, [m.weight for m in mods]) and isinstance(act, (torch.nn.ELU, torch.nn.Identity)):
                return K.skinny_encoders(xs, [m.weight for m in mods], [m.bias for m in mods], elu=isinstance(act, torch.nn.ELU))
        sa = torch.cat((self.state_input_encoder(state), self.action_input_encoder(action)), dim=-1)
        return self.uni_model_input_mapping_activation_func(sa) if self.separate_encoder else sa

    def forward(self, state, lst_state, lst_action, action, rnn_memory: Optional[RNNHidden], reward, detach_embedding=False,
                embedded=None):
        """embedded: what `embed(...)` returned for the same (state, lst_state, lst_action, rnn_memory, reward) -- the
        context encoder does not depend on `action`, so a caller may run it ahead of (or beside) the policy that produces
        the action."""
        emb_in = None if embedded is not None else self.get_embedding_input(state, lst_state, lst_action, reward)
        value, rnn_memory, emb, full = self.meta_forward(emb_in, self.state_action(state, action), rnn_memory, detach_embedding,
                                                         embedded=embedded)
        return value, emb, rnn_memory, full

    def embed(self, state, lst_state, lst_action, rnn_memory: Optional[RNNHidden], reward):
        """The context-encoder half of forward(): (embedding, its hidden, its per-layer record)."""
        return self._meta_forward_embedding(self.get_embedding_input(state, lst_state, lst_action, reward), rnn_memory)

    def forward_embedding(self, state, lst_state, lst_action, rnn_memory, reward):
        return self.get_embedding(self.get_embedding_input(state, lst_state, lst_action, reward), rnn_memory)


class ContextualTD3Value(ContextualSACValue):
    def __init__(self, *args, **kwargs):
        kwargs.setdefault('name', 'ContextualTD3Value')
        super().__init__(*args, **kwargs)
