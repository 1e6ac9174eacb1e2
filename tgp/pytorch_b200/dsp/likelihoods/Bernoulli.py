"""Bernoulli likelihood with a probit link (reference code/dsp/likelihoods/Bernoulli.py:21-157)."""
import torch
import torch.nn as nn
import torch.distributions as td

from .. import config as cg
from ..quadrature import GaussHermiteQuadrature1D
from . import _rows
from ... import functional as Fn


class Bernoulli(nn.Module):
    def __init__(self):
        super().__init__()
        self.C = 2
        self.quad_points = cg.quad_points
        self.quadrature_distribution = GaussHermiteQuadrature1D(self.quad_points)
        self.link_function = td.normal.Normal(0, 1).cdf

    def sample_from_output(self, f, i, **kwargs):
        return td.Bernoulli(probs=self.link_function(f)).sample().to(cg.dtype)

    def expected_log_prob_rows(self, Y, gauss_mean, gauss_cov, flow, X):
        assert len(flow) == 1, 'Flow list must be size 1 for Bernoulli likelihood'
        assert gauss_mean.size(0) == 1, 'Binary classification just require one GP for both classes'
        assert len(X.shape) == 3, 'Bad input X, expected (n_class,MB*S,Dx)'
        y = Y.reshape(-1).to(torch.float64)
        # the kernel clamps v < 0 to 0 and zeroes its gradient (Bernoulli.py:77)
        s, r = _rows.expected_log_prob_rows('bernoulli', self.quad_points, y, gauss_mean[0], gauss_cov[0], None, flow[0], X[0])
        return s, r.unsqueeze(0)

    def expected_log_prob(self, Y, gauss_mean, gauss_cov, flow, X, **kwargs):
        """int q(f) log p(y | Phi(G(f))) df summed over the minibatch (a scalar, as in the reference)."""
        return self.expected_log_prob_rows(Y, gauss_mean, gauss_cov, flow, X)[0]

    def marginal_moments(self, gauss_mean, gauss_cov, flow, X, **kwargs):
        """P(y=1|x), shape (MB, 1).  Identity flow: closed form (Rasmussen & Williams 3.80).  Otherwise quadrature with
        the batch-wide std of the variances, `gauss_cov.std()` — a defect of the reference (Bernoulli.py:120,141) kept
        deliberately so that results match it; see DESIGN.md "reference defects"."""
        assert len(flow) == 1, 'Flow list must be size 1 for Bernoulli likelihood'
        assert gauss_mean.size(0) == 1, 'Binary classification just require one GP for both classes'
        from ..models.flow import CompositeFlow, IdentityFlow      # here: `models` imports this package
        fl = flow[0]
        subs = fl.flow_arr if isinstance(fl, CompositeFlow) else [fl]
        identity = all(isinstance(f, IdentityFlow) for f in subs)
        bern_std = None if identity else Fn.global_std(gauss_cov).reshape(1).to(torch.float64).contiguous()   # all ranks' rows
        _, P, _ = _rows.test_rows('bernoulli', self.quad_points, None, gauss_mean[0], gauss_cov[0], None,
                                  None if identity else fl, X[0], bern_std=bern_std)
        return P.unsqueeze(1)
