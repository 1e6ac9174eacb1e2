from .utils_models import instance_kernel, return_mean, enable_eval_dropout
from .sparse_MF_SP import sparse_MF_SP
from .sparse_MF_GP import sparse_MF_GP

__all__ = ['instance_kernel', 'return_mean', 'enable_eval_dropout', 'sparse_MF_SP', 'sparse_MF_GP']
