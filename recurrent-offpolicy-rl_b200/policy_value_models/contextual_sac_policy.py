"""Tanh-Gaussian recurrent policy (single head: the universal network emits [logstd | mean]).

API-compatible with ContextualSACPolicySingleHead / ContextualSACPolicy
(ref: offpolicy_rnn/policy_value_models/contextual_sac_policy_single_head.py:11-129,
contextual_sac_policy.py:4-15): same constructor kwargs, same registered module names
(`state_encoder`, `last_act_encoder`, `last_obs_encoder`, `reward_encoder`), same forward tuple
`(action_mean, embedding, action_sample, log_prob, rnn_memory, full_rnn_memory)`, logstd is the FIRST
half of the head output (ref :105), clamp [-20, 2], softplus-corrected tanh log-density (ref :118-120).
`noise_fn` is the only addition: the standard-normal draw is injectable so that runs can be compared.
"""
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

from .. import kernels as K
from ..models.contextual_model import ContextualModel
from ..models.linear import Linear
from ..models.RNNHidden import RNNHidden
from .utils import nearest_power_of_two, nearest_power_of_two_half


class _InputEncoders:
    """The separate 128-d linear encoders of state / last state / last action / reward shared by policy
    and value (ref: contextual_sac_policy_single_head.py:33-52, contextual_sac_value.py:25-47)."""

    def _build_encoders(self, state_dim, action_dim, reward_input, last_action_input, last_state_input, separate_encoder):
        self.reward_input, self.last_action_input, self.last_state_input = reward_input, last_action_input, last_state_input
        self.reward_dim = 1 if reward_input else 0
        self.last_act_dim = action_dim if last_action_input else 0
        self.last_obs_dim = state_dim if last_state_input else 0
        self.separate_encoder = separate_encoder
        if separate_encoder:
            w = 128
            self.state_encoder = Linear(state_dim, w)
            self.last_act_encoder = Linear(self.last_act_dim, w) if self.last_act_dim else None
            self.reward_encoder = Linear(self.reward_dim, w) if self.reward_dim else None
            self.last_obs_encoder = Linear(self.last_obs_dim, w) if self.last_obs_dim else None
            return w * (1 + sum(e is not None for e in (self.last_act_encoder, self.last_obs_encoder, self.reward_encoder)))
        self.state_encoder = self.last_act_encoder = self.reward_encoder = self.last_obs_encoder = torch.nn.Identity()
        return state_dim + self.reward_dim + self.last_act_dim + self.last_obs_dim

    def _register_encoders(self):
        if self.separate_encoder:
            self.contextual_register_rnn_base_module(self.state_encoder, 'state_encoder')
            for mod, name in ((self.last_act_encoder, 'last_act_encoder'), (self.last_obs_encoder, 'last_obs_encoder'),
                              (self.reward_encoder, 'reward_encoder')):
                if mod is not None:
                    self.contextual_register_rnn_base_module(mod, name)

    def get_embedding_input(self, state, lst_state, lst_action, reward) -> torch.Tensor:
        pairs = [(self.state_encoder, state)]
        if self.last_state_input:
            pairs.append((self.last_obs_encoder, lst_state))
        if self.last_action_input:
            pairs.append((self.last_act_encoder, lst_action))
        if self.reward_input:
            pairs.append((self.reward_encoder, reward))
        if self.separate_encoder:
            xs, Ws, bs = [x for _, x in pairs], [m.weight for m, _ in pairs], [m.bias for m, _ in pairs]
            if K.skinny_encoders_ok(xs, Ws):          # every encoder writes its column block of one buffer: no cat
                return K.skinny_encoders(xs, Ws, bs)
        return torch.cat([m(x) for m, x in pairs], dim=-1)


class ContextualSACPolicySingleHead(ContextualModel, _InputEncoders):
    MAX_LOG_STD = 2.0
    MIN_LOG_STD = -20.0

    def __init__(self, state_dim, action_dim, embedding_size, embedding_hidden, embedding_activations,
                 embedding_layer_type, uni_model_hidden, uni_model_activations, uni_model_layer_type, fix_rnn_length,
                 uni_model_input_mapping_dim: int = 0, reward_input=False, last_action_input=True, last_state_input=False,
                 separate_encoder=False, output_logstd=True, name='ContextualSACPolicy'):
        if uni_model_activations[-1] != 'linear':
            uni_model_activations = uni_model_activations[:-1] + ['linear']
        if uni_model_layer_type[-1] != 'fc':
            raise NotImplementedError(f'It is not supported to construct {uni_model_layer_type[-1]} logstd and mean head!')
        if embedding_size == 'auto':
            embedding_size = nearest_power_of_two_half(state_dim)
        if uni_model_input_mapping_dim == 'auto':
            uni_model_input_mapping_dim = nearest_power_of_two(state_dim)
        cum_dim = self._build_encoders(state_dim, action_dim, reward_input, last_action_input, last_state_input, separate_encoder)
        ContextualModel.__init__(self, embedding_input_size=cum_dim, embedding_size=embedding_size,
                                 embedding_hidden=embedding_hidden, embedding_activations=embedding_activations,
                                 embedding_layer_type=embedding_layer_type, uni_model_input_size=state_dim,
                                 uni_model_output_size=action_dim * 2 if output_logstd else action_dim,
                                 uni_model_hidden=uni_model_hidden, uni_model_activations=uni_model_activations,
                                 uni_model_layer_type=uni_model_layer_type, fix_rnn_length=fix_rnn_length,
                                 uni_model_input_mapping_dim=uni_model_input_mapping_dim,
                                 uni_model_input_mapping_activation=embedding_activations[-1], name=name)
        self._register_encoders()
        self.state_dim, self.action_dim = state_dim, action_dim
        self.soft_plus = torch.nn.Softplus()
        self.noise_fn = torch.randn_like

    def forward(self, state, lst_state, lst_action, rnn_memory: Optional[RNNHidden], reward=None, detach_embedding=False):
        emb_in = self.get_embedding_input(state, lst_state, lst_action, reward)
        out, rnn_memory, emb, full = self.meta_forward(emb_in, state, rnn_memory, detach_embedding)
        if out.is_cuda and out.dtype == torch.float32:
            # one kernel for the whole head (forward) and one for its backward; the draw stays injectable
            noise = self.noise_fn(out[..., self.action_dim:]).detach()
            action_mean, action_sample, log_prob = K.tanh_gaussian_head(out, noise, self.MIN_LOG_STD, self.MAX_LOG_STD)
        else:
            logstd, logit = out.chunk(2, dim=-1)
            action_mean, action_sample, log_prob = self.process_model_out(logit, logstd)
        return action_mean, emb, action_sample, log_prob, rnn_memory, full

    def process_model_out(self, logit, logstd):
        logstd = torch.clamp(logstd, self.MIN_LOG_STD, self.MAX_LOG_STD)
        noise = self.noise_fn(logit).detach()
        sample = logit + noise * logstd.exp()
        log_prob = (-0.5 * noise.pow(2) - (logstd + 0.5 * np.log(2 * np.pi))).sum(-1, keepdim=True)
        log_prob = log_prob - (2 * (-sample - F.softplus(-2 * sample) + np.log(2))).sum(-1, keepdim=True)
        return torch.tanh(logit), torch.tanh(sample), log_prob

    def forward_embedding(self, state, lst_state, lst_action, rnn_memory, reward):
        return self.get_embedding(self.get_embedding_input(state, lst_state, lst_action, reward), rnn_memory)


class ContextualSACPolicy(ContextualSACPolicySingleHead):
    pass
