"""ORACLE (test infrastructure) -- import the UNMODIFIED reference and drive its own `train_one_batch`.

The reference package is taken from /root/reference (build container) or, where that does not exist (the GPU box),
from the byte-identical staged copy `oracle/_ref/` made by `oracle/make_ref.py`.

Import recipe: `offpolicy_rnn/__init__.py` drags in gym and smart_logger, which are not installed,
so the top-level package is pre-registered as a bare namespace and only the sub-packages on the update path are
imported; `selective_scan_cuda` (binary-only, absent everywhere) is stubbed so that `selective_scan_interface_new`
imports.  With `gpu_semantics=True` the smamba GPU semantics are obtained by routing `selective_scan_fn` to the authors'
own `selective_scan_ref` and forcing `Mamba.forward` through `forward_sequential` (used for the golden fixtures); with
`gpu_semantics=False` the reference runs exactly as it would on a CPU-only machine (its per-step `Mamba.step` loop) --
that is the reference's own CPU implementation, the one `bench.py --impl reference` times.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_root():
    for root in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isdir(os.path.join(root, "offpolicy_rnn")):
            return root
    return None


def available() -> bool:
    return reference_root() is not None


def load_reference(gpu_semantics=True):
    root = reference_root()
    if root is None:
        raise RuntimeError("the reference is neither at /root/reference nor staged under oracle/_ref (run oracle/make_ref.py)")
    if "offpolicy_rnn" in sys.modules and getattr(sys.modules["offpolicy_rnn"], "_is_ref", False):
        pkg = sys.modules["offpolicy_rnn"]
    else:
        pkg = types.ModuleType("offpolicy_rnn")
        pkg.__path__ = [os.path.join(root, "offpolicy_rnn")]
        pkg._is_ref = True
        pkg._root = root
        sys.modules["offpolicy_rnn"] = pkg
        sys.modules.setdefault("selective_scan_cuda", types.ModuleType("selective_scan_cuda"))
        sl = types.ModuleType("smart_logger")
        sl.Logger = object
        sl.get_customized_value = lambda *a, **k: 1000
        sys.modules.setdefault("smart_logger", sl)
    import offpolicy_rnn.models.smamba.mamba as smamba
    from offpolicy_rnn.models.smamba.mamba_ssm.ops import selective_scan_interface_new as ssi
    if not hasattr(smamba.Mamba, "_orig_forward"):
        smamba.Mamba._orig_forward = smamba.Mamba.forward
        smamba._orig_selective_scan_fn = smamba.selective_scan_fn

    def _gpu_semantics_forward(self, x, hidden=None, rnn_start=None, mask=None):
        out = self.forward_sequential(x, mask, rnn_start)
        if hidden is None:
            import torch
            hidden = torch.zeros((1, x.shape[0], self.conv_hidden_dim + self.ssm_hidden_dim))
        return out, hidden

    if gpu_semantics:
        smamba.selective_scan_fn = ssi.selective_scan_ref
        smamba.Mamba.forward = _gpu_semantics_forward
    else:
        smamba.selective_scan_fn = smamba._orig_selective_scan_fn
        smamba.Mamba.forward = smamba.Mamba._orig_forward
    return pkg


def install_algo_stubs():
    """gym / smart_logger / envs stand-ins: only what the algorithm modules touch at import."""
    root = reference_root()
    gym = types.ModuleType("gym")
    gym.Env = object
    gym.Space = object
    gym.Wrapper = object
    spaces = types.ModuleType("gym.spaces")
    spaces.Box = type("Box", (), {})
    spaces.Discrete = type("Discrete", (), {})
    gym.spaces = spaces
    wr = types.ModuleType("gym.wrappers")
    wr.RescaleAction = object
    gym.wrappers = wr
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.wrappers": wr})
    sl = sys.modules["smart_logger"]
    sl.Logger = object
    sl.experiment_config = types.SimpleNamespace()
    sl.init_config = lambda *a, **k: None
    pt = types.ModuleType("smart_logger.parameter")
    ptt = types.ModuleType("smart_logger.parameter.ParameterTemplate")
    ptt.ParameterTemplate = object
    pt.ParameterTemplate = ptt
    sys.modules.update({"smart_logger.parameter": pt, "smart_logger.parameter.ParameterTemplate": ptt})
    envs = types.ModuleType("envs")
    mpe = types.ModuleType("envs.make_pomdp_env")
    mpe.make_pomdp_env = lambda *a, **k: None
    pc = types.ModuleType("envs.pomdp_config")
    pc.env_config = {}
    sys.modules.update({"envs": envs, "envs.make_pomdp_env": mpe, "envs.pomdp_config": pc})
    if root not in sys.path:
        sys.path.insert(0, root)


REF_HP = dict(utd=1, policy_utd=1, randomize_mask=False, valid_number_post_randomized=0, random_trunc_traj=False,
              randomize_first_hidden=False, gamma=0.99, sac_tau=0.995, policy_update_per=1, no_alpha_auto_tune=False,
              policy_max_gradnorm=None, policy_embedding_max_gradnorm=None, value_max_gradnorm=None,
              value_embedding_max_gradnorm=None, redq_m=2, target_action_noise_std=0.04, target_action_noise_clip=0.12,
              policy_lr=3e-4, value_lr=1e-3, rnn_policy_lr=1e-5, rnn_value_lr=1e-5, alpha_lr=1e-4, policy_l2_norm=0.0,
              value_l2_norm=0.0, sample_std=0.1, target_entropy_ratio=1.0)


def algo_class(cls_name):
    import importlib
    mods = {"SACFullLengthRNNREDQ_SEP_OPTIM": "sac_full_length_rnn_redq_sep_optim",
            "TD3FullLengthRNNREDQ_SEP_OPTIM": "td3_full_length_rnn_redq_sep_optim",
            "SACFullLengthRNNEnsembleQ": "sac_full_length_rnn_ensembleQ",
            "TD3FullLengthRNNEnsembleQ": "td3_full_length_rnn_ensembleQ",
            "SACFullLengthRNNREDQ": "sac_full_length_rnn_redq",
            "TD3FullLengthRNNREDQ": "td3_full_length_rnn_redq",
            "SACFullLengthRNNENSEMBLEQ_SEP_OPTIM": "sac_full_length_rnn_ensembleQ_sep_optim"}
    return getattr(importlib.import_module("offpolicy_rnn.algorithm." + mods[cls_name]), cls_name)


def build_algorithm(cls_name, hp, policy_kwargs, value_kwargs, max_traj_len, act_dim, device="cpu", perturb=0.0, discrete=False):
    """The reference algorithm object, constructed without its environment / logger (`object.__new__` + exactly the
    attributes SAC.__init__ and the subclass __init__s set; ref: algorithm/sac.py:34-127,
    sac_full_length_rnn_ensembleQ.py:17-55, sac_full_length_rnn_redq_sep_optim.py:81-102).  `hp` must already hold
    `sac_batch_size`.  Returns the object with an EMPTY replay buffer (fill it with mem_push)."""
    import torch
    from offpolicy_rnn.algorithm.sac_full_length_rnn_redq_sep_optim import prepare_param_list
    from offpolicy_rnn.buffers.transition_buffer.nested_replay_memory import NestedMemoryArray
    from offpolicy_rnn.policy_value_models.make_models import make_policy_model, make_value_model
    from offpolicy_rnn.utility.q_value_guard import QValueGuard
    from offpolicy_rnn.utility.timer import Timer
    cls = algo_class(cls_name)
    algo = "td3" if cls_name.startswith("TD3") else "sac"
    sep = cls_name.endswith("SEP_OPTIM")
    A = object.__new__(cls)
    hp = dict(hp)
    if algo == "td3" or discrete:                      # ref: sac.py:72-74, td3_full_length_rnn_ensembleQ.py:21
        hp["no_alpha_auto_tune"] = True
    A.parameter = types.SimpleNamespace(**hp)
    A.timer = Timer()
    A.device = A.sample_device = torch.device(device)
    A.discrete_env = bool(discrete)
    A.base_algorithm = algo
    A.logger = lambda *a, **k: None
    pk = dict(policy_kwargs)
    if algo == "td3":
        pk["sample_std"] = hp["sample_std"]
    A.policy_args, A.value_args = pk, dict(value_kwargs)
    A.policy = make_policy_model(pk, algo, discrete)
    A.values = [make_value_model(value_kwargs, algo, discrete)]
    A.target_values = [make_value_model(value_kwargs, algo, discrete)]
    for m in [A.policy] + A.values + A.target_values:
        m.to(A.device)
    if perturb:
        for m in [A.policy] + A.values:
            with torch.no_grad():
                for p in m.parameters():
                    if p.dim() == 1 or p.abs().max() == 0:
                        p.add_(perturb * torch.randn_like(p))
    A._value_update(tau=0.0)
    import math
    a0 = math.log(hp.get("sac_alpha", 1.0)) if hp.get("no_alpha_auto_tune") else 0.0                       # ref: sac.py:75-78
    A.log_sac_alpha = torch.full((1,), a0, requires_grad=True, device=A.device)
    A.target_entropy = hp["target_entropy_ratio"] if discrete else -float(act_dim) * hp["target_entropy_ratio"]   # ref: sac.py:80
    for net in (A.values[0].embedding_network.layer_list + A.target_values[0].embedding_network.layer_list
                + A.values[0].uni_network.layer_list + A.target_values[0].uni_network.layer_list):
        if hasattr(net, 'desire_ndim'):
            net.desire_ndim = 4
        if hasattr(net, 'in_proj') and hasattr(net.in_proj, 'desire_ndim'):
            net.in_proj.desire_ndim = 4
    A.amp_scalar = A.amp_scalar_critic = None
    if cls._get_whether_require_amp(A):
        from torch.cuda.amp import GradScaler
        A.amp_scalar, A.amp_scalar_critic = GradScaler(), GradScaler()
    A.Q_guard = QValueGuard(True, True, 1.0 if discrete else 1 - 1e-3)                                      # ref: sac_full_length_rnn_ensembleQ.py:43-46
    A.target_policy = make_policy_model(pk, algo, discrete)
    A.target_policy.to(A.device)
    A.target_policy.copy_weight_from(A.policy, tau=0.0)
    A.target_policy.eval()
    A.optim_class = torch.optim.AdamW
    if sep:     # ref: sac_full_length_rnn_redq_sep_optim.py:85-92
        A.optimizer_policy = torch.optim.AdamW(prepare_param_list(A.policy, hp["rnn_policy_lr"], hp["policy_l2_norm"]),
                                               lr=hp["policy_lr"], weight_decay=hp["policy_l2_norm"])
        A.optimizer_value = torch.optim.AdamW(prepare_param_list(A.values[0], hp["rnn_value_lr"], hp["value_l2_norm"]),
                                              lr=hp["value_lr"], weight_decay=hp["value_l2_norm"])
    else:       # ref: sac.py:81-90
        A.optimizer_policy = torch.optim.AdamW(A.policy.parameters(True), lr=hp["policy_lr"], weight_decay=hp["policy_l2_norm"])
        A.optimizer_value = torch.optim.AdamW([p for v in A.values for p in v.parameters(True)], lr=hp["value_lr"],
                                              weight_decay=hp["value_l2_norm"])
    A.value_parameters = [p for v in A.values for p in v.parameters(True)]
    A.value_embedding_parameters = [p for v in A.values for p in v.embedding_network.parameters(True)]
    A.optimizer_alpha = torch.optim.AdamW([A.log_sac_alpha], lr=hp["alpha_lr"])
    A.grad_num = 0
    A.allow_nest_stack = cls.allow_nest_stack_trajs(A)
    A.replay_buffer = NestedMemoryArray(hp.get("max_buffer_transition_num", 1000), max_traj_len,
                                        additional_history_len=cls._get_skip_len(A))          # ref: sac_full_length_rnn_ensembleQ.py:41
    for v in A.values:
        v.train()
    for v in A.target_values:
        v.eval()
    A.policy.train()
    return A
