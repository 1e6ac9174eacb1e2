"""p(y | G(f)) Gaussian with a flow-warped mean (reference code/dsp/likelihoods/GaussianNonLinearMean.py:20-203)."""
import torch
import torch.nn as nn
import torch.distributions as td

from .. import config as cg
from ..quadrature import GaussHermiteQuadrature1D
from ..utils import positive_transform, inverse_positive_transform
from . import _rows


class GaussianNonLinearMean(nn.Module):
    def __init__(self, out_dim, noise_init, noise_is_shared, quadrature_points):
        super().__init__()
        self.out_dim = out_dim
        self.noise_is_shared = noise_is_shared
        init = inverse_positive_transform(torch.tensor(noise_init, dtype=cg.dtype))
        self.log_var_noise = nn.Parameter(torch.ones(1 if noise_is_shared else out_dim, 1, dtype=cg.dtype) * init)
        self.quad_points = quadrature_points
        self.quadrature_distribution = GaussHermiteQuadrature1D(quadrature_points)

    def _noise(self):
        return self.log_var_noise.expand(self.out_dim, 1) if self.noise_is_shared else self.log_var_noise

    def sample_from_output(self, f, i, **kwargs):
        var = positive_transform(self._noise()[i])
        return td.Normal(f, torch.ones_like(f) * torch.sqrt(var)).sample()

    def expected_log_prob_rows(self, Y, gauss_mean, gauss_cov, flow, X):
        """(ELL (Dy,), per-row terms (Dy, MB)) — the fused epilogue, one launch per output."""
        assert len(flow) == self.out_dim, 'The number of callables representing non linearities is different from out_dim'
        assert len(X.shape) == 3, 'Bad input X, expected (out_dim,MB*S,Dx)'
        assert X.size(0) == self.out_dim, 'Wrong first dimension in X, expected out_dim'
        if cg.positive_transform != 'exp':
            raise NotImplementedError("the fused epilogue implements positive_transform='exp' (the reference default)")
        noise = self._noise()
        sums, rows = [], []
        for dy in range(self.out_dim):
            s, r = _rows.expected_log_prob_rows('gauss_nonlinear', self.quad_points, Y[dy], gauss_mean[dy], gauss_cov[dy],
                                                noise[dy], flow[dy], X[dy])
            sums.append(s)
            rows.append(r)
        return torch.stack(sums), torch.stack(rows)

    def expected_log_prob(self, Y, gauss_mean, gauss_cov, flow, X, **kwargs):
        """E_q(f)[log p(y|G(f))] by Gauss-Hermite quadrature, reduced over the minibatch: shape (Dy,)."""
        return self.expected_log_prob_rows(Y, gauss_mean, gauss_cov, flow, X)[0]

    def marginal_moments(self, gauss_mean, gauss_cov, flow, X, **kwargs):
        """First moment and variance of p(y|x) = int p(y|G(f)) N(f|mean,cov) df: shapes (Dy, MB)."""
        assert len(X.shape) == 3, 'Bad input X, expected (out_dim,MB*S,Dx)'
        assert X.size(0) == self.out_dim, 'Wrong first dimension in X, expected out_dim'
        m1, m2 = [], []
        # NOTE the reference does not expand a shared noise here (GaussianNonLinearMean.py:177); with
        # noise_is_shared=False (every shipped configuration) both readings coincide
        noise = self._noise()
        for dy in range(self.out_dim):
            _, a, b = _rows.test_rows('gauss_nonlinear', self.quad_points, None, gauss_mean[dy], gauss_cov[dy], noise[dy],
                                      flow[dy], X[dy])
            m1.append(a)
            m2.append(b)
        return torch.stack(m1), torch.stack(m2)
