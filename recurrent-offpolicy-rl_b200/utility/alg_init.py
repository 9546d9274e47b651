"""Algorithm-name switch for the full-length recurrent algorithms
(ref: offpolicy_rnn/utility/alg_init.py:16-47; the MLP / sliced-sequence entries are out of scope)."""


def alg_class(name: str):
    from ..algorithm.sac_full_length_rnn_ensembleQ import SACFullLengthRNNEnsembleQ
    from ..algorithm.sac_full_length_rnn_redq import SACFullLengthRNNREDQ
    from ..algorithm.sac_full_length_rnn_redq_sep_optim import SACFullLengthRNNREDQ_SEP_OPTIM
    from ..algorithm.sac_full_length_rnn_ensembleQ_sep_optim import SACFullLengthRNNENSEMBLEQ_SEP_OPTIM
    from ..algorithm.td3_full_length_rnn_ensembleQ import TD3FullLengthRNNEnsembleQ
    from ..algorithm.td3_full_length_rnn_redq import TD3FullLengthRNNREDQ
    from ..algorithm.td3_full_length_rnn_redq_sep_optim import TD3FullLengthRNNREDQ_SEP_OPTIM
    table = {
        'sac_rnn_full_horizon_ensembleQ': SACFullLengthRNNEnsembleQ,
        'sac_rnn_full_horizon_redQ': SACFullLengthRNNREDQ,
        'sac_rnn_full_horizon_redQ_sep_optim': SACFullLengthRNNREDQ_SEP_OPTIM,
        'sac_rnn_full_horizon_ensemble_q_sep_optim': SACFullLengthRNNENSEMBLEQ_SEP_OPTIM,
        'td3_rnn_full_horizon_ensembleQ': TD3FullLengthRNNEnsembleQ,
        'td3_rnn_full_horizon_redQ': TD3FullLengthRNNREDQ,
        'td3_rnn_full_horizon_redQ_sep_optim': TD3FullLengthRNNREDQ_SEP_OPTIM,
    }
    if name not in table:
        raise NotImplementedError(f'{name}: only the full-length recurrent algorithms are on the hot path')
    return table[name]
