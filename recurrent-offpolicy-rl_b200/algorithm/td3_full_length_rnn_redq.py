"""TD3 + REDQ (ref: offpolicy_rnn/algorithm/td3_full_length_rnn_redq.py:10-50)."""
from .full_length_update import FullLengthRNNUpdate


class TD3FullLengthRNNREDQ(FullLengthRNNUpdate):
    base_algorithm = 'td3'
    use_redq = True
