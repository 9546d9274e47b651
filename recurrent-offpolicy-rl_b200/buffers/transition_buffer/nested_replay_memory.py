"""Nested trajectory sampler: padded, nest-stacked `[rows, Lmax, F]` batches with pre-step targets, start
flags and the valid indicator (layout: SURVEY.md App. A).

Same interface and bit-exact output as the reference's NestedMemoryArray.sample_trajs
(ref: offpolicy_rnn/buffers/transition_buffer/nested_replay_memory.py:8-25 row width, :38-56
`load_equalize` best-fit packing, :103-185 batch assembly), split B200-style:

  host   `plan_trajs()`   -- RNG draws (same numpy global-RNG call order), lengths, bin packing: integers only
  device `rorl_traj_gather` -- all data movement, from the device-resident fp32 ring (csrc/gather.cu)

`sample_trajs_device()` is what the update uses; `sample_trajs()` keeps the reference's numpy return
type by copying that device batch back.  `randomize_mask` / `random_trunc_traj` are
not part of the update path and are rejected.
"""
from __future__ import annotations

import math
from typing import List, NamedTuple, Tuple

import numpy as np
import torch

from .replay_memory import MemoryArray, Transition, tuplenames
from ... import _native as N


class SamplePlan(NamedTuple):
    entries: np.ndarray      # int64 [ntraj, 4] = (src_start, row, ptr, len) per placed trajectory
    row_end: np.ndarray      # int64 [rows] first step of each row's all-start tail
    rows: int
    width: int               # returned row width = max(ptr_end) + 1, clipped to the row length
    total_size: int          # sum of valid steps
    lens: np.ndarray         # float64 [rows, kmax]: leading 1, then len_i + skip per trajectory (traj_len_array)


class NestedMemoryArray(MemoryArray):
    def __init__(self, max_transition_num: int = 1000, max_traj_step: int = 1000, rnn_slice_length: int = 1,
                 additional_history_len: int = 0, map_to_two_power=True, device=None):
        row_len = max_traj_step + 2 + additional_history_len
        if map_to_two_power:
            row_len = self.nearest_power_of_two(row_len)
        super().__init__(max_transition_num, row_len, rnn_slice_length, device=device)
        self._additional_history_len = additional_history_len
        self._skip_step = 1 + additional_history_len
        self._colmap_dev = None
        self._batch_cache = None

    @staticmethod
    def nearest_power_of_two(x):
        e = int(math.ceil(math.log(x, 2)))
        return int(math.ceil(2 ** max(e, 0)))

    def load_equalize(self, traj_lens, max_traj_length) -> List[List[int]]:
        """Best-fit packing in arrival order; a trajectory fits only if strictly shorter than the room left
        (ref :46 uses `>`), and argmin breaks ties towards the earliest row."""
        bins: List[List[int]] = []
        room: List[int] = []
        for idx in range(len(traj_lens)):
            n = traj_lens[idx]
            if bins:
                left = [r - n if r > n else max_traj_length + 1 for r in room]
                best = int(np.argmin(left))
                if left[best] <= max_traj_length:
                    bins[best].append(idx)
                    room[best] = left[best]
                    continue
            bins.append([idx])
            room.append(max_traj_length - n)
        return bins

    def _init_memory_buffer(self, transition: Transition):
        super()._init_memory_buffer(transition)
        r = self.name2range
        self._source_range = r['state'] + r['reward_input'] + r['last_state']
        self._target_range = r['next_state'] + r['reward'] + r['state']
        self._action_range = r['action']
        self._mask_range = r['mask']
        self._rnn_start_range = r['start']

    # ---- host: integer plan -----------------------------------------------------------------------------------
    def plan_trajs(self, batch_size, max_sample_size=None, get_all=False, nest_stack_trajs=True) -> SamplePlan:
        inds = np.arange(self.available_traj_num) if get_all else self._traj_ind_sample(batch_size, max_sample_size)
        skip = self._skip_step
        lens = [self.trajectory_length[i] + skip for i in inds]
        starts = [self.trajectory_start[i] for i in inds]
        groups = self.load_equalize(lens, self.max_traj_step) if nest_stack_trajs else [[i] for i in range(len(lens))]
        entries, row_end, summary, width = [], [], [], 0
        for r, grp in enumerate(groups):
            p, ll = 0, [1]
            for k in grp:
                entries.append((starts[k], r, p, lens[k] - skip))
                ll.append(lens[k])
                p += lens[k]
            width = max(width, p)
            row_end.append(p)
            summary.append(ll)
        width = min(width + 1, self.max_traj_step)
        lens_arr = np.zeros((len(groups), max(len(s) for s in summary)))
        for r, s in enumerate(summary):
            lens_arr[r, :len(s)] = s
        return SamplePlan(np.asarray(entries, dtype=np.int64).reshape(-1, 4), np.asarray(row_end, dtype=np.int64),
                          len(groups), width, int(sum(lens) - len(lens) * skip), lens_arr)

    # ---- device: gather -----------------------------------------------------------------------------------------
    def _colmap(self):
        if self._colmap_dev is None:
            pairs = [c for dst, src in zip(self._target_range, self._source_range) for c in (dst, src)]
            cm = [self._rnn_start_range[0], self._mask_range[0], len(self._target_range)] + pairs
            self._colmap_dev = torch.tensor(cm, dtype=torch.int32, device=self.device)
        return self._colmap_dev

    def gather_device(self, plan: SamplePlan) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.device_buffer is None:
            raise RuntimeError('NestedMemoryArray was built without a CUDA device: the trajectory gather has no CPU path')
        F = self.device_buffer.shape[1]
        rows, Lmax = plan.rows, self.max_traj_step
        if self._batch_cache is None or self._batch_cache[0].shape[0] < rows:
            self._batch_cache = (torch.empty((rows, Lmax, F), dtype=torch.float32, device=self.device),
                                 torch.empty((rows, Lmax), dtype=torch.float32, device=self.device))
        batch, valid = self._batch_cache[0][:rows], self._batch_cache[1][:rows]
        host_plan = torch.from_numpy(np.concatenate((plan.entries.reshape(-1), plan.row_end))).pin_memory()
        dplan = host_plan.to(self.device, non_blocking=True)
        N.call("rorl_traj_gather", N.ptr(self.device_buffer), F, N.ptr(dplan), plan.entries.shape[0], N.ptr(self._colmap()),
               self._rnn_start_range[0], N.ptr(batch), N.ptr(valid), rows, Lmax, self._skip_step,
               int(plan.entries[:, 3].max()), N.stream())
        dplan.record_stream(torch.cuda.current_stream())
        return batch[:, :plan.width], valid[:, :plan.width].unsqueeze(-1)

    def _reject_unsupported(self, randomize_mask, random_trunc_traj):
        if randomize_mask or random_trunc_traj:
            raise NotImplementedError('randomize_mask / random_trunc_traj are not part of the update hot path')

    def sample_trajs_device(self, batch_size, max_sample_size=None, get_all=False, randomize_mask=False,
                            valid_number_post_randomized=0, equalize_data_of_each_traj=False, random_trunc_traj=False,
                            copy=False, nest_stack_trajs=True):
        """-> (Transition of device fp32 views, total_size, valid indicator [rows, W, 1], traj_len_array)."""
        self._reject_unsupported(randomize_mask, random_trunc_traj)
        plan = self.plan_trajs(batch_size, max_sample_size, get_all, nest_stack_trajs)
        batch, valid = self.gather_device(plan)
        return self.array_to_transition(batch), plan.total_size, valid, plan.lens

    def sample_trajs(self, batch_size, max_sample_size=None, get_all=False, randomize_mask=False,
                     valid_number_post_randomized=0, equalize_data_of_each_traj=False, random_trunc_traj=False,
                     copy=False, nest_stack_trajs=True):
        """Reference-typed result (numpy float64 views), produced by the same device gather."""
        tr, total, valid, lens = self.sample_trajs_device(batch_size, max_sample_size, get_all, randomize_mask,
                                                         valid_number_post_randomized, equalize_data_of_each_traj,
                                                         random_trunc_traj, copy, nest_stack_trajs)
        to_np = lambda t: None if t is None else t.cpu().numpy().astype(np.float64)
        return Transition(*[to_np(f) for f in tr]), total, to_np(valid), lens
