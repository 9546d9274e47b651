"""RESeL: SAC + REDQ with the context-encoder-specific learning rate
(ref: offpolicy_rnn/algorithm/sac_full_length_rnn_redq_sep_optim.py:37-107).  The per-encoder split is always
on in FullLengthRNNUpdate (set rnn_*_lr == *_lr to disable it), so this class only fixes the flags."""
from .full_length_update import prepare_param_list  # noqa: F401  (re-exported: same name as the reference helper)
from .sac_full_length_rnn_redq import SACFullLengthRNNREDQ


class SACFullLengthRNNREDQ_SEP_OPTIM(SACFullLengthRNNREDQ):
    pass
