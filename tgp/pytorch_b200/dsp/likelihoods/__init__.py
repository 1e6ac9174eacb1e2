from .GaussianLinearMean import GaussianLinearMean
from .GaussianNonLinearMean import GaussianNonLinearMean
from .Bernoulli import Bernoulli

__all__ = ['GaussianLinearMean', 'GaussianNonLinearMean', 'Bernoulli']
