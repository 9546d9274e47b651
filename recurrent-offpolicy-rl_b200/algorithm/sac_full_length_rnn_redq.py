"""SAC + REDQ: target = min over a random `redq_m` subset, actor = mean over the ensemble
(ref: offpolicy_rnn/algorithm/sac_full_length_rnn_redq.py:10-47)."""
from .full_length_update import FullLengthRNNUpdate


class SACFullLengthRNNREDQ(FullLengthRNNUpdate):
    base_algorithm = 'sac'
    use_redq = True
