"""Defaults for `init_params` of the model constructors (what the reference keeps in code/dsp/models/config_models.py):
the variational distribution q(u) = N(mean_scale * 1, variance_scale * I) unless the caller overrides it, as the
reference's main.py does with variance_scale = 1e-5 (main.py:104-110)."""
import copy

_DEFAULTS = {'variational_distribution': {'mean_scale': 0.0, 'variance_scale': 1.0}}


def get_init_params(params):
    """Fills the keys the caller did not provide; the caller's dict is updated in place and returned."""
    for name, default in _DEFAULTS.items():
        if name not in params:
            params[name] = copy.deepcopy(default)
    return params
