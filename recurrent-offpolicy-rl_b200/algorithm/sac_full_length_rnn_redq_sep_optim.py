"""RESeL: SAC + REDQ with the context-encoder-specific learning rate
(ref: offpolicy_rnn/algorithm/sac_full_length_rnn_redq_sep_optim.py:37-107).  `sep_optim` makes the shared engine
build the per-encoder AdamW groups of `prepare_param_list`; without it one AdamW runs over all parameters at
policy_lr / value_lr, as in the reference's plain classes (ref: sac.py:81-90)."""
from .full_length_update import prepare_param_list  # noqa: F401  (re-exported: same name as the reference helper)
from .sac_full_length_rnn_redq import SACFullLengthRNNREDQ


class SACFullLengthRNNREDQ_SEP_OPTIM(SACFullLengthRNNREDQ):
    sep_optim = True
