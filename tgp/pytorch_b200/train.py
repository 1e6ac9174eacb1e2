"""CUDA-graph training step: the reference's inner loop

    loss = -model.ELBO(x, y)[0]; optimizer.zero_grad(); loss.backward(); optimizer.step()      (trainer_base.py:329-342)

captured ONCE as a CUDA graph (our kernels, the Adam update and the few torch glue kernels — about 170 launches) and
replayed per minibatch.  At the reference's own problem sizes (boston: 455 rows, M = 100) the step is launch-latency
bound, which is exactly what a graph removes; the three `.item()` host syncs per step of the reference trainer
(trainers_regression.py:88-106) become one optional read-back every `k` steps.

Constraints of capture: build the step BEFORE any eager `backward()` of the same model on the default stream (autograd
ties each parameter's gradient accumulator to the stream of its first backward); fixed minibatch shape (the last, shorter batch of an epoch runs eagerly through `eager_step`),
no host synchronisation inside the step, hence `cg.check_cholesky_status` is switched off while capturing and the pivot
status of every engine is inspected after replay instead (`check()`), where the eager path would have raised.
"""
import torch

from .dsp import config as cg


class GraphedElboStep:
    def __init__(self, model, optimizer, x_example, y_example, warmup=3):
        if not x_example.is_cuda:
            raise RuntimeError('CUDA graphs need CUDA tensors (no CPU path)')
        for group in optimizer.param_groups:
            if not group.get('capturable', False):
                raise ValueError('build the optimizer with capturable=True (e.g. torch.optim.Adam(params, lr, capturable=True))')
        self.model, self.opt = model, optimizer
        self.x = x_example.clone()
        self.y = y_example.clone()
        self.loss = None
        self.ell = None
        self.kld = None
        self._old_check = cg.check_cholesky_status
        cg.check_cholesky_status = False
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):               # lazy one-time work (module loads, workspaces) happens here
                    self._step_body()
            torch.cuda.current_stream().wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            self.opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(self.graph):
                self._step_body()
        finally:
            cg.check_cholesky_status = self._old_check

    def _step_body(self):
        ELBO, ELL, KLD = self.model.ELBO(self.x, self.y)
        loss = -ELBO
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        self.loss, self.ell, self.kld = loss.detach(), ELL.detach(), KLD.detach()

    def __call__(self, x, y):
        """One training step on a minibatch of the captured shape; returns the (device-resident) loss tensor."""
        if x.shape != self.x.shape or y.shape != self.y.shape:
            raise ValueError('minibatch shape differs from the captured one; use eager_step for ragged batches')
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        # the replayed optimiser step changed the parameters in place without touching their `_version`: drop any
        # factorisation cached for evaluation batches (functional._QfMarginals keys it on identity + version)
        self.model._invalidate_eval_cache()
        return self.loss

    def eager_step(self, x, y):
        ELBO, _, _ = self.model.ELBO(x, y)
        loss = -ELBO
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        self.model._invalidate_eval_cache()
        return loss.detach()

    def check(self):
        """Raises if any factorisation inside the replayed steps hit a non-positive pivot (the eager path would have
        taken psd_safe_cholesky's jitter ladder there)."""
        for eng in self.model._engines.values():
            bad = int(eng.status.item())
            if bad:
                raise RuntimeError('cholesky: non-positive pivot %d inside a graphed step; rerun the step eagerly '
                                   '(GraphedElboStep.eager_step) to take the jitter ladder' % bad)
