"""ORACLE (test infrastructure) -- numpy restatement of the reference's transition ring buffer and
nested trajectory sampler.  Bit-exact contract, including the order in which the numpy GLOBAL RNG
is consumed.

ref: offpolicy_rnn/buffers/transition_buffer/replay_memory.py (MemoryArray) and
     offpolicy_rnn/buffers/transition_buffer/nested_replay_memory.py (NestedMemoryArray).
"""
from __future__ import annotations

import math
from collections import namedtuple

import numpy as np

FIELDS = ('state', 'last_state', 'last_action', 'action', 'next_state', 'reward', 'logp', 'mask', 'start', 'done',
          'reward_input', 'timeout')  # ref: replay_memory.py:11
Transition = namedtuple('Transition', FIELDS)


def pow2_ceil(x):
    """ref: nested_replay_memory.py:27-36"""
    e = int(math.ceil(math.log(x, 2)))
    return int(math.ceil(2 ** max(e, 0)))


def field_width(item):
    """ref: replay_memory.py:160-170"""
    if isinstance(item, np.ndarray):
        return item.shape[-1]
    if isinstance(item, list):
        return len(item)
    if item is None:
        return 0
    return 1


class RefNestedReplay:
    def __init__(self, max_transition_num, max_traj_step, additional_history_len=0):
        # ref: nested_replay_memory.py:9-25
        self.row_len = pow2_ceil(max_traj_step + 2 + additional_history_len)
        self.capacity = int(max_transition_num)
        self.skip = 1 + additional_history_len
        self.buf = None
        self.cols = None
        self.traj_len, self.traj_start = [], []
        self.ptr = 0
        self.count = 0
        self.pending = []
        self.cache = None

    # -- storage (ref: replay_memory.py:119-234) -------------------------------------------------
    def _init(self, tr):
        self.cols, c = {}, 0
        for name, item in zip(FIELDS, tr):
            w = field_width(item)
            self.cols[name] = list(range(c, c + w))
            c += w
        self.buf = np.zeros((self.capacity + self.row_len, c))
        self.src_cols = self.cols['state'] + self.cols['reward_input'] + self.cols['last_state']   # :70
        self.dst_cols = self.cols['next_state'] + self.cols['reward'] + self.cols['state']         # :71

    def _row(self, tr):
        parts = []
        for item in tr:
            if isinstance(item, np.ndarray):
                parts.append(item.reshape((1, -1)))
            elif isinstance(item, list):
                parts.append(np.array(item).reshape((1, -1)))
            elif item is None:
                continue
            else:
                parts.append(np.array([[item]]))
        return np.hstack(parts)

    def mem_push(self, tr):
        self.pending.append(tr)
        if np.all(tr.done):
            if np.all(tr.mask):
                self._complete(self.pending)
            self.pending = []

    def _complete(self, traj):
        if self.buf is None:
            self._init(traj[0])
        n, drop = len(traj), 0
        if self.count + n > self.capacity:
            c = self.count
            while c + n > self.capacity:
                c -= self.traj_len[drop]
                drop += 1
        if drop:
            self.count -= sum(self.traj_len[:drop])
            del self.traj_start[:drop]
            del self.traj_len[:drop]
        self.traj_start.append(self.ptr)
        for tr in traj:
            self.buf[self.ptr] = 0
            self.buf[self.ptr, :] = self._row(tr)
            self.ptr += 1
        self.traj_len.append(n)
        self.count += n
        if self.ptr >= self.capacity:
            self.ptr = 0

    # -- index draw (ref: replay_memory.py:56-90) ------------------------------------------------
    def _draw(self, batch_size):
        ntraj = len(self.traj_len)
        want = int(np.ceil(batch_size / (self.count / ntraj)))
        perm = np.random.permutation(ntraj)
        if want <= ntraj:
            inds = perm[:want]
        else:
            inds = np.random.randint(0, ntraj, (want,))
        total = sum(self.traj_len[i] for i in inds)
        extra = []
        while total < batch_size:
            pos = want + len(extra)
            if ntraj > pos:
                i = perm[pos]
            else:
                i = np.random.randint(low=0, high=ntraj)
            total += self.traj_len[i]
            extra.append(i)
        if extra:
            inds = np.concatenate((inds, np.array(extra)), axis=0)
        return inds

    # -- bin packing (ref: nested_replay_memory.py:38-56) ------------------------------------------
    @staticmethod
    def pack(lens, cap):
        bins, room = [], []
        for i, n in enumerate(lens):
            if bins:
                left = [r - n if r > n else cap + 1 for r in room]
                j = int(np.argmin(left))
                if left[j] <= cap:
                    bins[j].append(i)
                    room[j] = left[j]
                    continue
            bins.append([i])
            room.append(cap - n)
        return bins

    # -- ref: nested_replay_memory.py:84-100 ----------------------------------------------------------
    @staticmethod
    def equalized_valid_nums(lens_plus_skip, desired_total):
        order = np.argsort(lens_plus_skip)
        n = len(lens_plus_skip)
        avg = int(np.ceil(desired_total / n))
        out, got = [avg] * n, 0
        for i in range(n):
            tl = lens_plus_skip[order[i]] - 1
            want = int(np.ceil((desired_total - got) / (n - i)))
            if want <= 0:
                want = avg
            if want > tl:
                want = tl
            got += want
            out[order[i]] = want
        return out

    # -- ref: nested_replay_memory.py:103-185 -----------------------------------------------------------
    def sample_trajs(self, batch_size, nest_stack_trajs=True, randomize_mask=False, valid_number_post_randomized=0,
                     equalize_data_of_each_traj=True, random_trunc_traj=False):
        if random_trunc_traj:
            batch_size *= 2                                                              # :109-110
        inds = self._draw(batch_size)
        if random_trunc_traj:
            lens = [np.random.randint(0, self.traj_len[i]) + 1 + self.skip for i in inds]   # :114
        else:
            lens = [self.traj_len[i] + self.skip for i in inds]
        starts = [self.traj_start[i] for i in inds]
        eq = randomize_mask and equalize_data_of_each_traj
        if eq:
            valid_nums = self.equalized_valid_nums(lens, valid_number_post_randomized)     # :116-117
        groups = self.pack(lens, self.row_len) if nest_stack_trajs else [[i] for i in range(len(lens))]
        rows = len(groups)
        total = int(sum(lens) - len(lens) * self.skip)
        F = self.buf.shape[-1]
        if self.cache is None or self.cache.shape[0] < rows:
            self.cache = np.zeros((rows, self.row_len, F))
        else:
            self.cache[:rows] = 0
        g = self.cache
        mcol, scol = self.cols['mask'][0], self.cols['start'][0]
        valid = g[:, :, mcol:mcol + 1].copy()
        summary, width = [], 0
        for r, grp in enumerate(groups):
            p, ll = 0, [1]
            for k in grp:
                n, s0 = lens[k], starts[k]
                ll.append(n)
                g[r, p + self.skip:p + n, :] = self.buf[s0:s0 + (n - self.skip), :]
                g[r, p + self.skip - 1, self.dst_cols] = self.buf[s0, self.src_cols]
                g[r, p + self.skip - 1, self.cols['action']] = 0
                g[r, p:p + self.skip, scol] = 1
                valid[r, p + self.skip:p + n, :] = self.buf[s0:s0 + (n - self.skip), mcol:mcol + 1]
                if eq:                                                                   # :166-168
                    zeros = np.random.permutation(n - self.skip)[:-valid_nums[k]] + p + self.skip
                    g[r, zeros, mcol] = 0
                p += n
            width = max(width, p)
            g[r, p:, scol] = 1
            summary.append(np.array(ll))
        width += 1
        lens_arr = np.zeros((rows, max(len(s) for s in summary)))
        for r, s in enumerate(summary):
            lens_arr[r, :len(s)] = s
        data = g[:rows, :width, :]
        fields = [data[..., c[0]:c[-1] + 1] if len(c) else None for c in (self.cols[n] for n in FIELDS)]
        tr = Transition(*fields)
        if randomize_mask and not equalize_data_of_each_traj:
            # _mask_rnd_select (:78-82, called :183-184): `mask.reshape((-1,))` of the sliced batch is a COPY unless the
            # slice spans the whole row length, so the zeroing only lands in the batch in that case; the permutation
            # draw happens either way (and advances the global RNG)
            m1 = tr.mask.reshape((-1,))
            idx = m1.nonzero()[0]
            kill = idx[np.random.permutation(idx.shape[0])[:-valid_number_post_randomized]]
            m1[kill] = 0
        return tr, total, valid[:rows, :width, :], lens_arr
