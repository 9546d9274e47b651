"""SAC, full-length trajectories, min over the whole Q ensemble for target and actor
(ref: offpolicy_rnn/algorithm/sac_full_length_rnn_ensembleQ.py:16-132)."""
from .full_length_update import FullLengthRNNUpdate


class SACFullLengthRNNEnsembleQ(FullLengthRNNUpdate):
    base_algorithm = 'sac'
    use_redq = False
