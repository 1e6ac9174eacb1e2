"""tgp.pytorch_b200 — B200-native minibatch ELBO / test-NLL path for SVGP, TGP and ID_TGP.

`tgp.pytorch_b200.dsp` mirrors the class surface of the reference's `code/dsp` package; the arithmetic runs in
`libtgp_b200.so` (hand-written sm_100a kernels behind the C-ABI of include/tgp_b200.h).  No CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ['_lib']
