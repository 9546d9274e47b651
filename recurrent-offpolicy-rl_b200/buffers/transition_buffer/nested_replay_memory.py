"""Nested trajectory sampler: padded, nest-stacked `[rows, Lmax, F]` batches with pre-step targets, start
flags and the valid indicator (layout: SURVEY.md App. A).

Same interface and bit-exact output as the reference's NestedMemoryArray.sample_trajs
(ref: offpolicy_rnn/buffers/transition_buffer/nested_replay_memory.py:8-25 row width, :38-56
`load_equalize` best-fit packing, :103-185 batch assembly), split B200-style:

  host   `plan_trajs()`   -- RNG draws (same numpy global-RNG call order), lengths, bin packing: integers only
  device `rorl_traj_gather` -- all data movement, from the device-resident fp32 ring (csrc/gather.cu)

`sample_trajs_device()` is what the update uses; `sample_trajs()` keeps the reference's numpy return
type by copying that device batch back.  The sampler options `random_trunc_traj` and `randomize_mask` (equalised per
trajectory or global; ref :78-100,109-117,166-168,183-184) are part of the host plan too: the extra RNG draws happen
in the reference's order and the mask entries to clear are applied on the device after the gather.
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
    mask_zero: np.ndarray    # int64 [k]: row * Lmax + step of batch `mask` entries cleared by randomize_mask


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

    def get_equalized_valid_num_each_traj(self, traj_len_added_1, desired_total_valid_number):
        """Valid-step budget per trajectory, shortest first (ref :84-100; np.argsort's default ordering is part of it)."""
        order = np.argsort(traj_len_added_1)
        n = len(traj_len_added_1)
        avg = int(np.ceil(desired_total_valid_number / n))
        out, got = [avg for _ in range(n)], 0
        for i in range(n):
            tl = traj_len_added_1[order[i]] - 1
            want = int(np.ceil((desired_total_valid_number - got) / (n - i)))
            if want <= 0:
                want = avg
            if want > tl:
                want = tl
            got += want
            out[order[i]] = want
        return out

    # ---- host: integer plan -----------------------------------------------------------------------------------
    def plan_trajs(self, batch_size, max_sample_size=None, get_all=False, nest_stack_trajs=True, randomize_mask=False,
                   valid_number_post_randomized=0, equalize_data_of_each_traj=False, random_trunc_traj=False) -> SamplePlan:
        if get_all:
            inds = np.arange(self.available_traj_num)
        else:
            if random_trunc_traj:
                batch_size *= 2                                                          # ref :109-110
            inds = self._traj_ind_sample(batch_size, max_sample_size)
        skip = self._skip_step
        if random_trunc_traj:                                                            # ref :114: one draw per trajectory, in order
            lens = [np.random.randint(0, self.trajectory_length[i]) + 1 + skip for i in inds]
        else:
            lens = [self.trajectory_length[i] + skip for i in inds]
        starts = [self.trajectory_start[i] for i in inds]
        equalized = randomize_mask and equalize_data_of_each_traj
        if equalized:
            valid_nums = self.get_equalized_valid_num_each_traj(lens, valid_number_post_randomized)
        groups = self.load_equalize(lens, self.max_traj_step) if nest_stack_trajs else [[i] for i in range(len(lens))]
        Lmax = self.max_traj_step
        entries, row_end, summary, width, mask_zero = [], [], [], 0, []
        for r, grp in enumerate(groups):
            p, ll = 0, [1]
            for k in grp:
                entries.append((starts[k], r, p, lens[k] - skip))
                ll.append(lens[k])
                if equalized:                                                            # ref :166-168
                    zeros = np.random.permutation(lens[k] - skip)[:-valid_nums[k]] + p + skip
                    mask_zero.append(r * Lmax + zeros.astype(np.int64))
                p += lens[k]
            width = max(width, p)
            row_end.append(p)
            summary.append(ll)
        width = min(width + 1, self.max_traj_step)
        if randomize_mask and not equalize_data_of_each_traj:
            # _mask_rnd_select on the returned batch (ref :78-82,183-184).  The draw always happens; the zeroing only
            # lands in the batch when `mask.reshape((-1,))` of the [rows, width, 1] slice is a VIEW, i.e. when the
            # slice spans the whole row length or there is a single row (otherwise numpy hands back a copy).
            mcol = self._mask_range[0]
            nz = []
            for src, r, p, n in entries:
                t = np.nonzero(self.memory_buffer[src:src + n, mcol])[0] + p + skip
                nz.append(r * width + t[t < width])
            nz = np.sort(np.concatenate(nz)) if nz else np.zeros(0, dtype=np.int64)
            kill = nz[np.random.permutation(nz.shape[0])[:-valid_number_post_randomized]]
            if width == self.max_traj_step or len(groups) == 1:
                mask_zero.append((kill // width) * Lmax + kill % width)
        lens_arr = np.zeros((len(groups), max(len(s) for s in summary)))
        for r, s in enumerate(summary):
            lens_arr[r, :len(s)] = s
        mz = np.concatenate(mask_zero).astype(np.int64) if mask_zero else np.zeros(0, dtype=np.int64)
        return SamplePlan(np.asarray(entries, dtype=np.int64).reshape(-1, 4), np.asarray(row_end, dtype=np.int64),
                          len(groups), width, int(sum(lens) - len(lens) * skip), lens_arr, mz)

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
        if plan.mask_zero.size:                      # randomize_mask: clear the chosen entries of the batch's mask column
            idx = torch.from_numpy(plan.mask_zero * F + self._mask_range[0]).pin_memory().to(self.device, non_blocking=True)
            self._batch_cache[0].view(-1).index_fill_(0, idx, 0.0)
        return batch[:, :plan.width], valid[:, :plan.width].unsqueeze(-1)

    def sample_trajs_device(self, batch_size, max_sample_size=None, get_all=False, randomize_mask=False,
                            valid_number_post_randomized=0, equalize_data_of_each_traj=False, random_trunc_traj=False,
                            copy=False, nest_stack_trajs=True):
        """-> (Transition of device fp32 views, total_size, valid indicator [rows, W, 1], traj_len_array)."""
        plan = self.plan_trajs(batch_size, max_sample_size, get_all, nest_stack_trajs, randomize_mask,
                               valid_number_post_randomized, equalize_data_of_each_traj, random_trunc_traj)
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
