"""SVGP (Hensman et al.) = the transformed model with an identity flow per output
(reference code/dsp/models/sparse_MF_GP.py:39-98)."""
from .sparse_MF_SP import sparse_MF_SP


class sparse_MF_GP(sparse_MF_SP):
    def __init__(self, model_specs, X, init_Z, N, likelihood, num_outputs, is_whiten, K_is_shared, mean_is_shared,
                 Z_is_shared, q_U_is_shared, add_noise_inducing, init_params={}):
        flow_specs = [[('identity', [])] for _ in range(num_outputs)]
        super().__init__(model_specs, X, init_Z, N, likelihood, num_outputs, is_whiten, K_is_shared, mean_is_shared,
                         Z_is_shared, q_U_is_shared, flow_specs, 'single', add_noise_inducing, be_fully_bayesian=False,
                         init_params=init_params)

    def sample_from_variational_marginal(self, X, S, diagonal, is_duvenaud, init_Z=None):
        f, mean_q_f, cov_q_f, _ = super().sample_from_variational_marginal(X, S, diagonal, is_duvenaud, init_Z)
        return f, mean_q_f, cov_q_f, f
