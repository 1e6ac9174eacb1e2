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
import ctypes as C

import torch

from . import _lib
from .dsp import config as cg


class FusedAdam:
    """torch.optim.Adam's update (what the reference trains with: trainer_base.py:342 through optimizers.py:10-22) for ALL
    parameter tensors in ONE kernel launch (`tgp_adam_step`), step count on the device — a TGP with a StepTanhL(15,4) flow has
    ~280 parameter tensors, i.e. hundreds of element-wise launches per step through the stock optimiser.

    Same constructor shape as torch optimisers (an iterable of tensors or of param-group dicts with `lr` / `weight_decay`, as
    `Trainer_base` builds them: trainer_base.py:106-186).  Gradients live in persistent buffers (`p.grad` is set once and zeroed
    in place by `zero_grad`), so the device-side address table stays valid; FP64 CUDA parameters only."""

    def __init__(self, params, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        groups = list(params)
        if groups and not isinstance(groups[0], dict):
            groups = [{'params': groups}]
        self.param_groups = []
        for g in groups:
            full = dict(lr=lr, weight_decay=weight_decay, betas=betas, eps=eps, capturable=True)
            full.update(g)
            full['params'] = list(full['params'])
            self.param_groups.append(full)
        self.betas, self.eps = betas, eps
        ps = [p for g in self.param_groups for p in g['params']]
        if not ps:
            raise ValueError('optimizer got an empty parameter list')
        dev = ps[0].device
        for p in ps:
            if not (p.is_cuda and p.dtype == torch.float64 and p.is_contiguous()):
                raise ValueError('FusedAdam updates contiguous float64 CUDA parameters (no CPU path)')
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        self.state = {p: dict(exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p)) for p in ps}
        self.lib, self.device, self._ps = _lib.load(), dev, ps
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self._build_tables()

    def _build_tables(self):
        ps, dev = self._ps, self.device
        self._grads = [p.grad for p in ps]                      # keep the buffers whose addresses are in the table alive
        tab = [[p.data_ptr(), p.grad.data_ptr(), self.state[p]['exp_avg'].data_ptr(), self.state[p]['exp_avg_sq'].data_ptr()] for p in ps]
        lr, wd = [], []
        for g in self.param_groups:
            lr += [float(g['lr'])] * len(g['params'])
            wd += [float(g['weight_decay'])] * len(g['params'])
        bt, bo = [], []
        for i, p in enumerate(ps):
            for off in range(0, p.numel(), 4096):
                bt.append(i)
                bo.append(off)
        self.t_tab = torch.tensor(tab, dtype=torch.int64, device=dev)
        self.t_sizes = torch.tensor([p.numel() for p in ps], dtype=torch.int64, device=dev)
        self.t_lr = torch.tensor(lr, dtype=torch.float64, device=dev)
        self.t_wd = torch.tensor(wd, dtype=torch.float64, device=dev)
        self.t_bt = torch.tensor(bt, dtype=torch.int32, device=dev)
        self.t_bo = torch.tensor(bo, dtype=torch.int64, device=dev)
        self._lr_seen = lr

    def zero_grad(self, set_to_none=False):
        torch._foreach_zero_(self._grads)                       # in place: the gradient buffers keep their addresses

    def step(self):
        for p, g in zip(self._ps, self._grads):
            if p.grad is not g:                                  # somebody replaced a gradient tensor: rebuild the table
                if p.grad is None:
                    p.grad = g.zero_()
                self._build_tables()
                break
        lr = [float(g['lr']) for g in self.param_groups for _ in g['params']]
        if lr != self._lr_seen:                                  # a scheduler changed a learning rate
            self.t_lr.copy_(torch.tensor(lr, dtype=torch.float64))
            self._lr_seen = lr
        with torch.cuda.device(self.device):
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(self.lib.tgp_adam_step(len(self._ps), int(self.t_bt.numel()), self.t_tab.data_ptr(), self.t_sizes.data_ptr(),
                                              self.t_lr.data_ptr(), self.t_wd.data_ptr(), self.t_bt.data_ptr(), self.t_bo.data_ptr(),
                                              float(self.betas[0]), float(self.betas[1]), float(self.eps), self.step_dev.data_ptr(), st),
                       'tgp_adam_step')


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
