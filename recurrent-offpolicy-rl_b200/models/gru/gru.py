"""`gru` encoder layer: torch.nn.GRU(batch_first=True) parameters and call contract
(ref: offpolicy_rnn/models/rnn_base.py:59,245-247,454).  Subclassing nn.GRU keeps the reference's
state_dict keys (weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0) so checkpoints interchange; the
arithmetic does not go through cuDNN: the input projection is one tensor-core GEMM over all steps
and the L dependent steps run in the persistent cluster kernel of csrc/gru.cu.
"""
import torch.nn as nn

from ... import kernels as K


class GRULayer(nn.GRU):
    def __init__(self, input_size, hidden_size, batch_first=True):
        assert batch_first, "the encoder stack is batch-first (ref: rnn_base.py:245-247)"
        super().__init__(input_size, hidden_size, batch_first=batch_first)

    def forward(self, x, hx=None):
        """x [B, L, I]; hx [1, B, H] -> (out [B, L, H], h_n [1, B, H]), as torch.nn.GRU returns them."""
        if not x.is_cuda:
            raise RuntimeError("rorl_b200 GRU runs on the sm_100a persistent kernel only; there is no CPU path")
        gi = K.linear(x, self.weight_ih_l0, self.bias_ih_l0)
        out, h_last = K.gru_scan(gi, self.weight_hh_l0, self.bias_hh_l0, None if hx is None else hx[0])
        return out, h_last.unsqueeze(0)
