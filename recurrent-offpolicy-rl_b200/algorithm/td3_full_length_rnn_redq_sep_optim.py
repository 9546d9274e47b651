"""TD3 + REDQ with the RESeL learning-rate split
(ref: offpolicy_rnn/algorithm/td3_full_length_rnn_redq_sep_optim.py:37-95)."""
from .full_length_update import prepare_param_list  # noqa: F401
from .td3_full_length_rnn_redq import TD3FullLengthRNNREDQ


class TD3FullLengthRNNREDQ_SEP_OPTIM(TD3FullLengthRNNREDQ):
    sep_optim = True
