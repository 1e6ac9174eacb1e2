import torch
import torch.nn as nn


class CholeskyVariationalDistribution(nn.Module):
    """Holds `variational_mean` (*batch, M) and `chol_variational_covar` (*batch, M, M)."""

    def __init__(self, num_inducing_points, batch_shape=torch.Size([]), mean_init_std=1e-3, **kwargs):
        super().__init__()
        self.variational_mean = nn.Parameter(torch.zeros(*batch_shape, num_inducing_points))
        self.chol_variational_covar = nn.Parameter(torch.eye(num_inducing_points).repeat(*batch_shape, 1, 1))
