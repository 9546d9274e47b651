"""SAC with the min over the whole Q ensemble plus the RESeL learning-rate split
(ref: offpolicy_rnn/algorithm/sac_full_length_rnn_ensembleQ_sep_optim.py:21-82)."""
from .full_length_update import prepare_param_list  # noqa: F401
from .sac_full_length_rnn_ensembleQ import SACFullLengthRNNEnsembleQ


class SACFullLengthRNNENSEMBLEQ_SEP_OPTIM(SACFullLengthRNNEnsembleQ):
    sep_optim = True
