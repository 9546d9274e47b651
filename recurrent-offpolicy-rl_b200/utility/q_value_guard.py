"""Q-target clamp guard with its state on the device.

Same semantics as the reference's QValueGuard (ref: offpolicy_rnn/utility/q_value_guard.py:4-45): the first
`clamp` initialises min/max from the tensor itself, `update` widens them with the new extrema and then decays
them towards those extrema with ratio `decay_ratio`.  The reference keeps min/max as Python floats obtained
with 4 `.item()` syncs per update; here they live in a device double[4] = {min, max, initialised, decay} that
the fused target kernels read and write (csrc/losses.cu), and `get_min/get_max` read them back on demand.
"""
import torch


class QValueGuard:
    def __init__(self, guard_min=True, guard_max=True, decay_ratio=1.0, device=None):
        assert guard_min and guard_max, 'the update path constructs the guard with both bounds (ref: sac_full_length_rnn_ensembleQ.py:43-46)'
        self._decay_ratio = decay_ratio
        self.device = device
        self.state = None
        if device is not None:
            self.to(device)

    def to(self, device):
        self.device = device
        self.state = torch.tensor([1000000.0, -1000000.0, 0.0, self._decay_ratio], dtype=torch.float64, device=device)
        return self

    def reset(self):
        self.to(self.device)

    def get_min(self) -> float:
        return float(self.state[0].item())

    def get_max(self) -> float:
        return float(self.state[1].item())
