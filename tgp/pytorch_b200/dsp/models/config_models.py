"""Default initial parameters of the variational distribution (reference code/dsp/models/config_models.py)."""
init_params = {'variational_distribution': {'mean_scale': 0.0, 'variance_scale': 1.0}}


def get_init_params(params):
    for key, val in init_params.items():
        params.setdefault(key, val)
    return params
