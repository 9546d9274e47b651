"""`gru` encoder layer: torch.nn.GRU(batch_first=True) parameters and call contract
(ref: offpolicy_rnn/models/rnn_base.py:59,245-247,454).  Subclassing nn.GRU keeps the reference's
state_dict keys (weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0) so checkpoints interchange.
"""
import torch.nn as nn


class GRULayer(nn.GRU):
    def __init__(self, input_size, hidden_size, batch_first=True):
        super().__init__(input_size, hidden_size, batch_first=batch_first)
