"""rorl_b200 -- B200-native recurrent off-policy (RESeL) update hot path.

Host side mirrors the reference's interfaces for this path (layer registry and layer-ID strings,
policy_value_models, buffers, the full-length SAC/TD3 update and the RESeL optimizer split); the
arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI in include/rorl_b200.h.
Import as `rorl_b200` (see the shim at the repo root).
"""
__version__ = "0.1.0"
