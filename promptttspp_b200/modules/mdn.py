"""Mixture-density head parameter holder (reference: promptttspp/modules/mdn.py:37-78)."""
from torch import nn


class MDNLayer(nn.Module):
    def __init__(self, in_dim, out_dim, num_gaussians=30, dim_wise=False):
        super().__init__()
        self.in_dim, self.out_dim = in_dim, out_dim
        self.num_gaussians, self.dim_wise = num_gaussians, dim_wise
        self.log_pi = nn.Linear(in_dim, out_dim * num_gaussians if dim_wise else num_gaussians)
        self.log_sigma = nn.Linear(in_dim, out_dim * num_gaussians)
        self.mu = nn.Linear(in_dim, out_dim * num_gaussians)
