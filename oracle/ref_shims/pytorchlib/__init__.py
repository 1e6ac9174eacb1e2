"""Stand-in for jmaronas/pytorch_library@version-1.5.0 (source unavailable offline).

Layer order Linear -> [BatchNorm1d] -> activation -> [Dropout] is an ASSUMPTION (parity unpinned,
SURVEY.md §8c).  The dropout module must have 'Dropout' in its class name
(reference code/dsp/models/utils_models.py:358-364).
"""
import torch
import torch.nn as nn


def return_activation(name):
    return {'relu': nn.ReLU, 'tanh': nn.Tanh, 'linear': nn.Identity, 'sigmoid': nn.Sigmoid}[name]()


class apply_linear(nn.Module):
    def __init__(self, in_dim, out_dim, act, shape=None, std=0.0, drop=0.0, bn=0):
        super().__init__()
        layers = [nn.Linear(in_dim, out_dim)]
        if bn:
            layers.append(nn.BatchNorm1d(out_dim))
        layers.append(return_activation(act))
        if drop > 0:
            layers.append(nn.Dropout(drop))
        self.forward_lin = nn.Sequential(*layers)

    def forward(self, x):
        return self.forward_lin(x)


def compute_calibration_measures(probs, labels, apply_softmax=False, bins=15):
    if apply_softmax:
        probs = torch.softmax(probs, dim=1)
    nll = -torch.log(probs.gather(1, labels.view(-1, 1).long())).mean()
    return None, None, None, nll
