"""Import the UNMODIFIED reference (/root/reference) in THIS container only.

Used exclusively by tests/golden/make_golden.py to produce the committed
fixtures.  Nothing under tests/ that runs on the GPU box imports this file:
/root/reference does not exist there.

Recipe: `offpolicy_rnn/__init__.py` drags in gym and
smart_logger, which are not installed, so the top-level package is
pre-registered as a bare namespace and only the sub-packages on the update
path are imported.  `selective_scan_cuda` (binary-only, absent) is stubbed so
that `selective_scan_interface_new` imports; the smamba GPU semantics are then
obtained by routing `selective_scan_fn` to the authors' own
`selective_scan_ref` and forcing `Mamba.forward` through `forward_sequential`.
"""
import sys
import types

REF_ROOT = "/root/reference"


def load_reference():
    if "offpolicy_rnn" in sys.modules and getattr(sys.modules["offpolicy_rnn"], "_is_ref", False):
        return sys.modules["offpolicy_rnn"]
    pkg = types.ModuleType("offpolicy_rnn")
    pkg.__path__ = [REF_ROOT + "/offpolicy_rnn"]
    pkg._is_ref = True
    sys.modules["offpolicy_rnn"] = pkg
    sys.modules.setdefault("selective_scan_cuda", types.ModuleType("selective_scan_cuda"))
    # smart_logger / gym stubs (only what the algorithm modules touch at import)
    sl = types.ModuleType("smart_logger")
    sl.Logger = object
    sl.get_customized_value = lambda *a, **k: 1000
    sys.modules.setdefault("smart_logger", sl)

    import offpolicy_rnn.models.smamba.mamba as smamba
    from offpolicy_rnn.models.smamba.mamba_ssm.ops import selective_scan_interface_new as ssi

    smamba.selective_scan_fn = ssi.selective_scan_ref

    def _gpu_semantics_forward(self, x, hidden=None, rnn_start=None, mask=None):
        out = self.forward_sequential(x, mask, rnn_start)
        if hidden is None:
            import torch
            hidden = torch.zeros((1, x.shape[0], self.conv_hidden_dim + self.ssm_hidden_dim))
        return out, hidden

    smamba.Mamba.forward = _gpu_semantics_forward
    return pkg
