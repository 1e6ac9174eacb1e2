"""Host-side mirror of the reference's `code/dsp` package for the minibatch-ELBO / test-NLL path.

Same class, function and attribute names as the reference (so its `main.py` / `exp_utils.py` imports resolve:
`instance_kernel`, `sparse_MF_SP`, `sparse_MF_GP`, `instance_flow`, `SAL`, `StepTanhL`, `GaussianNonLinearMean`,
`GaussianLinearMean`, `Bernoulli`, `KMEANS` — reference code/main.py:26-37, code/exp_utils.py:7-8), but the
arithmetic of ELBO / marginals / expected log-lik / test log-lik is enqueued on the B200 through libtgp_b200.so.
"""
