import torch


def psd_safe_cholesky(A, upper=False, out=None, jitter=None):
    L = torch.linalg.cholesky(A)
    return L.mT if upper else L
