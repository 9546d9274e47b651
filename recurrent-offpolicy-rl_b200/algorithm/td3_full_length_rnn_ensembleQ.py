"""TD3, full-length trajectories, frozen target policy for the target action
(ref: offpolicy_rnn/algorithm/td3_full_length_rnn_ensembleQ.py:18-71)."""
from .full_length_update import FullLengthRNNUpdate


class TD3FullLengthRNNEnsembleQ(FullLengthRNNUpdate):
    base_algorithm = 'td3'
    use_redq = False
