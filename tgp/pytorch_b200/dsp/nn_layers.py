"""One dense layer of the input-dependent flow MLPs — stand-in for `pytorchlib.apply_linear`
(git+git://github.com/jmaronas/pytorch_library.git@version-1.5.0, not installable offline; reference call sites
code/dsp/models/flow.py:666-689, 859-868).

Layer order Linear -> [BatchNorm1d] -> activation -> [Dropout] is an assumption (parity unpinned at this boundary,
SURVEY.md §8c).  The dropout module keeps 'Dropout' in its class name because `enable_eval_dropout`
(reference code/dsp/models/utils_models.py:358-364) finds it by that substring.
"""
import torch.nn as nn

_ACT = {'relu': nn.ReLU, 'tanh': nn.Tanh, 'linear': nn.Identity, 'sigmoid': nn.Sigmoid}


def return_activation(name):
    return _ACT[name]()


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
