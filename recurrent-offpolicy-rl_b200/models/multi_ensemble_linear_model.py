"""`multi_num` x `num_ensemble` independent linear maps (the input projections of the ensemble encoders `egilr-E`,
`elru-E`).  Same parameters (`weight [multi, E, in, out]`, `bias [multi, E, 1, out]`) and rank dispatch as the reference
(ref: offpolicy_rnn/models/multi_ensemble_linear_model.py:9-74); on CUDA the contraction runs on the grouped tcgen05
GEMM (kernels.ensemble_linear) with the (multi, E) pairs as the group axis."""
import torch
import torch.nn as nn

from .. import kernels as K


class MultiEnsembleLinear(nn.Module):
    def __init__(self, input_dim: int, output_dim: int, num_ensemble: int, multi_num: int, bias: bool = True, desire_ndim: int = None):
        super().__init__()
        self.use_bias, self.desire_ndim, self.num_ensemble, self.multi_num = bias, desire_ndim, num_ensemble, multi_num
        self.weight = nn.Parameter(torch.zeros(multi_num, num_ensemble, input_dim, output_dim))
        if bias:
            self.bias = nn.Parameter(torch.zeros(multi_num, num_ensemble, 1, output_dim))
        nn.init.trunc_normal_(self.weight, std=1 / (2 * input_dim ** 0.5))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        assert self.desire_ndim is not None, 'desire_ndim: 3 for [B, C] data, 4 for [B, L, C] data'
        W = self.weight
        Mn, E, Kin, Nout = W.shape
        b = self.bias if self.use_bias else None
        nd = x.dim()
        if self.desire_ndim == 4 and x.is_cuda and nd in (3, 4, 5):
            Wg = W.reshape(Mn * E, Kin, Nout)
            bg = None if b is None else b.reshape(Mn * E, 1, Nout)
            y = None
            if nd == 3:                                   # shared [B, L, in] -> [multi, E, B, L, out]
                y = K.ensemble_linear(x, Wg, bg, False, shared=True)
            elif nd == 5:                                 # per (multi, member) input
                y = K.ensemble_linear(x.reshape(Mn * E, *x.shape[2:]), Wg, bg, False, shared=False)
            elif nd == 4:                                 # per-member input shared by the `multi` maps
                ys = [K.ensemble_linear(x, W[c], None if b is None else b[c], False, shared=False) for c in range(Mn)]
                if all(t is not None for t in ys):
                    return torch.stack(ys, dim=0)
            if y is not None:
                return y.reshape(Mn, E, *y.shape[1:])
        if nd == 2:
            assert self.desire_ndim == 3
            y = torch.einsum('ij,cbjk->cbik', x, W)
        elif nd == 3:
            y = torch.einsum('bij,cbjk->cbik', x, W) if self.desire_ndim == 3 else torch.einsum('hij,cbjk->cbhik', x, W)
        elif nd == 4:
            assert self.desire_ndim == 4
            y = torch.einsum('bhij,cbjk->cbhik', x, W)
        elif nd == 5:
            assert self.desire_ndim == 4
            y = torch.einsum('cbhij,cbjk->cbhik', x, W)
        else:
            raise NotImplementedError
        if b is not None:
            y = y + (b if self.desire_ndim == 3 else b.unsqueeze(2))
        return y
