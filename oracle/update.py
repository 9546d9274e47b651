"""ORACLE (test infrastructure) -- CPU restatement of one full-length recurrent SAC / TD3 update
(`train_one_batch` with utd = 1) on top of oracle.model, oracle.sampler and torch.optim.AdamW.

ref: offpolicy_rnn/algorithm/sac_full_length_rnn_ensembleQ.py:297-467 (train_one_batch), :83-132
     (_target_Q/_Q_loss/_policy_loss/_alpha_loss), sac_full_length_rnn_redq.py:16-47 (REDQ subset / mean),
     sac_full_length_rnn_redq_sep_optim.py:37-102 (RESeL param groups), td3_full_length_rnn_*.py,
     sac.py:61-95 (optimizers, log_alpha), rnn_base.py:475-491 (soft update).
"""
from __future__ import annotations

import copy
from typing import Callable, Dict

import numpy as np
import torch

from . import model as M
from .ops import QValueGuard


def clone_sd(sd, requires_grad=False):
    return {k: {n: t.detach().clone().requires_grad_(requires_grad) for n, t in v.items()} for k, v in sd.items()}


def param_groups(sd, rnn_lr, l2, sep_optim=True):
    """RESeL split by module NAME: the whole embedding_model goes to the slow group
    (ref: sac_full_length_rnn_redq_sep_optim.py:49-79); the plain classes run one AdamW group over all parameters
    (ref: sac.py:81-90)."""
    if not sep_optim:
        return [{'params': [t for mod in sd.values() for t in mod.values()]}]
    groups = []
    for k, mod in sd.items():
        params = list(mod.values())
        if k.endswith('encoder'):
            groups.append({'params': params})
        elif k == 'embedding_model':
            groups.append({'params': params, 'lr': rnn_lr, 'weight_decay': l2})
        else:
            groups.append({'params': params})
    return groups


class RefUpdate:
    """State of the reference algorithm object that `train_one_batch` touches."""

    def __init__(self, policy_sd, value_sd, policy_spec: M.ModelSpec, value_spec: M.ModelSpec, hp, replay,
                 noise_fn: Callable, algo='sac', redq=True, allow_nest_stack=True, module_order=None, sep_optim=True,
                 hidden_fn: Callable = None, discrete=False):
        self.hp = hp
        self.algo, self.redq = algo, redq
        self.pspec, self.vspec = policy_spec, value_spec
        self.policy = clone_sd(policy_sd, True)
        self.value = clone_sd(value_sd, True)
        self.target = clone_sd(value_sd, False)
        self.target_policy = clone_sd(policy_sd, False)
        import math
        self.discrete = discrete
        fixed_alpha = discrete or algo == 'td3' or hp.get('no_alpha_auto_tune', False)      # ref: sac.py:72-78
        self.log_alpha = torch.full((1,), math.log(hp.get('sac_alpha', 1.0)) if fixed_alpha else 0.0, requires_grad=True)
        self.target_entropy = (hp.get('target_entropy_ratio', 1.0) if discrete
                               else -float(policy_spec.action_dim) * hp.get('target_entropy_ratio', 1.0))
        self.opt_pi = torch.optim.AdamW(param_groups(self.policy, hp['rnn_policy_lr'], hp.get('policy_l2_norm', 0.0), sep_optim),
                                        lr=hp['policy_lr'], weight_decay=hp.get('policy_l2_norm', 0.0))
        self.opt_q = torch.optim.AdamW(param_groups(self.value, hp['rnn_value_lr'], hp.get('value_l2_norm', 0.0), sep_optim),
                                       lr=hp['value_lr'], weight_decay=hp.get('value_l2_norm', 0.0))
        self.hidden_fn = hidden_fn
        self.opt_alpha = torch.optim.AdamW([self.log_alpha], lr=hp['alpha_lr'])
        self.guard = QValueGuard(decay_ratio=1.0 if discrete else 1 - 1e-3)       # ref: sac_full_length_rnn_ensembleQ.py:43-46
        self.replay = replay
        self.noise_fn = noise_fn
        self.allow_nest_stack = allow_nest_stack
        self.grad_num = 0

    def _mask_mean(self, data, mask, n):
        return (data * mask).sum() / n

    def _one_update_discrete(self, did_policy, state, last_state, action, last_action, next_state, done, mask, reward, reward_input,
                             side_t, side_qt, side, side_pi, alpha, total):
        """Discrete action space.  ref: sac_full_length_rnn_ensembleQ.py:134-185 (+ REDQ: sac_full_length_rnn_redq.py:52-88);
        alpha is fixed (sac.py:72-74), so no alpha step."""
        hp = self.hp
        A = self.pspec.action_dim
        onehot = torch.nn.functional.one_hot(action.squeeze(-1).long(), num_classes=A).float()
        with torch.no_grad():
            lst_a = onehot if self.redq else action
            _, _, logp_next = M.policy_forward_discrete(self.policy, self.pspec, next_state, state, lst_a, side_t, reward)
            q_next, _ = M.value_forward_discrete(self.target, self.vspec, next_state, state, onehot, side_qt, reward)
            if self.redq:
                idx = np.random.permutation(q_next.shape[0])[:hp['redq_m']]
                q_next = q_next[idx, :]
            m = ((q_next.min(dim=0).values - alpha * logp_next) * logp_next.exp()).sum(dim=-1, keepdim=True)
            target_q = reward + (1 - done) * hp['gamma'] * self.guard.clamp(m)
        self.guard.update(target_q * mask)
        n_valid = mask.sum()
        q, _ = M.value_forward_discrete(self.value, self.vspec, state, last_state, last_action, side, reward_input)
        q_sel = q.gather(-1, action.long().unsqueeze(0).expand(q.shape[0], *action.shape))
        q_loss = sum(self._mask_mean((q_sel[i] - target_q).pow(2), mask, n_valid) for i in range(q_sel.shape[0]))
        self.opt_q.zero_grad()
        q_loss.backward()
        q_norm = self._clip(self.value, hp.get('value_max_gradnorm'), hp.get('value_embedding_max_gradnorm'))
        self.value_grads = {k: {n: (t.grad.clone() if t.grad is not None else None) for n, t in v.items()} for k, v in self.value.items()}
        self.opt_q.step()
        tau = hp['sac_tau']
        with torch.no_grad():
            for k in self.value:
                for n in self.value[k]:
                    tp = self.target[k][n]
                    tp.copy_(tp * tau + (1 - tau) * self.value[k][n])
        out = {'critic_loss': q_loss.item(), 'target_q_max': target_q.abs().max().item(), 'real_batch_size': total, 'value_grad_norm': q_norm}
        if did_policy:
            _, _, logp = M.policy_forward_discrete(self.policy, self.pspec, state, last_state, last_action, side_pi, reward_input)
            qp, _ = M.value_forward_discrete(self.value, self.vspec, state, last_state, last_action, side, reward_input, detach_embedding=True)
            agg = qp.mean(dim=0) if self.redq else qp.min(dim=0).values
            actor_loss = self._mask_mean((((alpha * logp) - agg) * logp.exp()).sum(dim=-1, keepdim=True), mask, n_valid)
            self.opt_pi.zero_grad()
            actor_loss.backward()
            out['policy_grad_norm'] = self._clip(self.policy, hp.get('policy_max_gradnorm'), hp.get('policy_embedding_max_gradnorm'))
            self.policy_grads = {k: {n: (t.grad.clone() if t.grad is not None else None) for n, t in v.items()} for k, v in self.policy.items()}
            self.opt_pi.step()
            out['actor_loss'] = actor_loss.item()
            out['log_prob'] = self._mask_mean((logp * logp.exp()).sum(dim=-1, keepdim=True), mask, n_valid).item()
            out['policy_l2_norm_square'] = sum(float((t.detach() ** 2).sum()) for k in ('embedding_model', 'universal_model', 'uni_input_mapping_network')
                                               if k in self.policy for t in self.policy[k].values())
        out['log_alpha'] = self.log_alpha.item()
        out['clip_min'], out['clip_max'] = self.guard.min, self.guard.max
        out['q1_l2_norm_square'] = sum(float((t.detach() ** 2).sum()) for k in ('embedding_model', 'universal_model', 'uni_input_mapping_network')
                                       if k in self.value for t in self.value[k].values())
        return out

    def _clip(self, sd, max_norm, emb_clip):
        """ref: sac_full_length_rnn_ensembleQ.py:239-250,274-287 -- global-norm clip, then value clip on the embedding
        network plus the hard-coded 1e-3 on every smamba A_log; returns the logged gradient norm."""
        norm = 0.0
        if max_norm is not None:
            norm = torch.nn.utils.clip_grad_norm_([t for m in sd.values() for t in m.values()], max_norm, norm_type=2).item()
        if emb_clip is not None:
            torch.nn.utils.clip_grad_value_(list(sd['embedding_model'].values()), emb_clip)
            a_logs = [t for n, t in sd['embedding_model'].items() if n.endswith('mixer.A_log') and '.layers.' in n]
            if a_logs:
                torch.nn.utils.clip_grad_value_(a_logs, 1e-3)
            norm = 0.0
        return norm

    def train_one_batch(self) -> Dict[str, float]:
        """utd iterations of `_one_update` with the reference's actor cadence (ref :311,405)."""
        hp = self.hp
        utd, cnt, out, pol = hp.get('utd', 1), 0, None, {}
        for i in range(utd):
            did = self.grad_num % hp.get('policy_update_per', 1) == 0 and (i + 1) / utd * hp.get('policy_utd', 1) > cnt
            out = self._one_update(did)
            cnt += int(did)
            pol.update({k: out[k] for k in ('actor_loss', 'alpha_loss', 'log_prob', 'policy_grad_norm', 'policy_l2_norm_square') if k in out})
        out.update(pol)
        self.grad_num += 1
        return out

    def _one_update(self, did_policy) -> Dict[str, float]:
        hp = self.hp
        f32 = lambda a: torch.from_numpy(a).to(torch.float32)
        batch, total, valid_ind, len_arr = self.replay.sample_trajs(hp['sac_batch_size'], nest_stack_trajs=self.allow_nest_stack)
        state, last_state, action, last_action, next_state, done, mask, reward, reward_input, timeout, rnn_start = (
            f32(getattr(batch, n)).clone() for n in ('state', 'last_state', 'action', 'last_action', 'next_state', 'done',
                                                      'mask', 'reward', 'reward_input', 'timeout', 'start'))
        valid_ind = f32(valid_ind).clone()
        total_start, total_valid = rnn_start.clone(), valid_ind.clone()                       # :338-341
        total_valid[torch.where(torch.diff(valid_ind, dim=-2) == 1)] = 1
        total_start[torch.where(torch.diff(total_start, dim=-2) == -1)] = 0
        done[timeout > 0] = 0                                                                 # :342
        alpha = self.log_alpha.exp().detach()
        att = f32(len_arr)                                                                    # :358-366
        att = torch.cat((att, torch.zeros((att.shape[0], state.shape[-2] - att.shape[1]))), dim=-1)
        tgt_att = torch.cat((att[..., 1:], torch.zeros((att.shape[0], 1))), dim=-1).to(torch.int)
        att = att.to(torch.int)
        h0_pi = h0_qt = h0_q = None
        if hp.get('randomize_first_hidden', False):       # ref :345-351 (one draw per model; both policy passes share theirs)
            h0_pi, h0_qt, h0_q = (self.hidden_fn(spec, state.shape[0]) for spec in (self.pspec, self.vspec, self.vspec))
        side_t = M.Side(total_start, total_valid, tgt_att, h0=h0_pi)
        side_qt = M.Side(total_start, total_valid, tgt_att, h0=h0_qt)
        side = M.Side(rnn_start, valid_ind, att, h0=h0_q)
        side_pi = M.Side(rnn_start, valid_ind, att, h0=h0_pi)
        td3 = self.algo == 'td3'
        if self.discrete:
            return self._one_update_discrete(did_policy, state, last_state, action, last_action, next_state, done, mask, reward,
                                             reward_input, side_t, side_qt, side, side_pi, alpha, total)
        # ---- target (no grad) -------------------------------------------------------------- :83-103
        with torch.no_grad():
            pol_t = self.policy if (self.redq or not td3) else self.target_policy
            noise = self.noise_fn(next_state.shape[:-1] + (self.pspec.action_dim,))
            a_mean, _, a_next, logp_next = M.policy_forward(pol_t, self.pspec, next_state, state, action, side_t, reward,
                                                            noise, td3, hp.get('sample_std', 0.1))
            if td3:
                n2 = self.noise_fn(a_mean.shape)
                a_next = torch.clamp(a_mean + torch.clamp(n2 * hp['target_action_noise_std'], -hp['target_action_noise_clip'],
                                                          hp['target_action_noise_clip']), -1, 1)
            q_next, _ = M.value_forward(self.target, self.vspec, next_state, state, action, a_next, side_qt, reward)
            if self.redq:
                idx = np.random.permutation(q_next.shape[0])[:hp['redq_m']]
                q_next = q_next[idx, :]
            m = q_next.min(dim=0).values
            if not td3:
                m = m - alpha * logp_next
            target_q = reward + (1 - done) * hp['gamma'] * self.guard.clamp(m)
        self.guard.update(target_q * mask)                                                    # :387
        n_valid = mask.sum()
        # ---- critic ------------------------------------------------------------------------ :105-114,261-295
        q, _ = M.value_forward(self.value, self.vspec, state, last_state, last_action, action, side, reward_input)
        q_loss = self._mask_mean((q - target_q.unsqueeze(0)).pow(2).sum(dim=0), mask, n_valid)
        self.opt_q.zero_grad()
        q_loss.backward()
        q_norm = self._clip(self.value, hp.get('value_max_gradnorm'), hp.get('value_embedding_max_gradnorm'))
        self.value_grads = {k: {n: (t.grad.clone() if t.grad is not None else None) for n, t in v.items()} for k, v in self.value.items()}
        self.opt_q.step()
        tau = hp['sac_tau']
        with torch.no_grad():                                                                 # rnn_base.py:483-491
            for k in self.value:
                for n in self.value[k]:
                    tp = self.target[k][n]
                    tp.copy_(tp * tau + (1 - tau) * self.value[k][n])
        out = {'critic_loss': q_loss.item(), 'target_q_max': target_q.abs().max().item(), 'real_batch_size': total,
               'value_grad_norm': q_norm}
        # ---- actor + alpha ----------------------------------------------------------------- :116-132,405-432
        if did_policy:
            noise = self.noise_fn(state.shape[:-1] + (self.pspec.action_dim,))
            a_mean, _, a_samp, logp = M.policy_forward(self.policy, self.pspec, state, last_state, last_action, side_pi,
                                                       reward_input, noise, td3, hp.get('sample_std', 0.1))
            a_in = a_mean if td3 else a_samp
            qp, _ = M.value_forward(self.value, self.vspec, state, last_state, last_action, a_in, side, reward_input,
                                    detach_embedding=True)
            agg = qp.mean(dim=0) if self.redq else qp.min(dim=0).values
            actor_loss = self._mask_mean((-agg) if td3 else (alpha * logp - agg), mask, n_valid)
            self.opt_pi.zero_grad()
            actor_loss.backward()
            out['policy_grad_norm'] = self._clip(self.policy, hp.get('policy_max_gradnorm'), hp.get('policy_embedding_max_gradnorm'))
            self.policy_grads = {k: {n: (t.grad.clone() if t.grad is not None else None) for n, t in v.items()} for k, v in self.policy.items()}
            self.opt_pi.step()
            out['actor_loss'] = actor_loss.item()
            out['policy_l2_norm_square'] = sum(float((t.detach() ** 2).sum()) for k in ('embedding_model', 'universal_model', 'uni_input_mapping_network')
                                               if k in self.policy for t in self.policy[k].values())
            if not td3:
                alpha_loss = -self._mask_mean(self.log_alpha * (logp + self.target_entropy).detach(), mask, n_valid)
                self.opt_alpha.zero_grad()
                alpha_loss.backward()
                self.opt_alpha.step()
                with torch.no_grad():
                    self.log_alpha.clamp_max_(1)
                out['alpha_loss'] = alpha_loss.item()
                out['log_prob'] = self._mask_mean(logp, mask, n_valid).item()
        out['log_alpha'] = self.log_alpha.item()
        out['clip_min'], out['clip_max'] = self.guard.min, self.guard.max
        out['q1_l2_norm_square'] = sum(float((t.detach() ** 2).sum()) for k in ('embedding_model', 'universal_model', 'uni_input_mapping_network')
                                       if k in self.value for t in self.value[k].values())   # ref :454, contextual_model.py:227-228
        return out
