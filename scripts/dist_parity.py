"""Multi-GPU parity (run under torchrun): every rank evaluates its row slice of a golden fixture's minibatch through the
class API; the all-reduced ELBO and gradients must equal the reference's single-process values."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from tests.golden_util import Golden, rel_err
from tests.model_util import build_from_golden
from tgp.pytorch_b200 import dist as D

warnings.simplefilter('ignore')
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = 'cuda:%d' % local
dist.init_process_group('nccl', device_id=torch.device(dev))
worst = 0.0
for name in ('synth_reg_d8_m64_p1', 'boston_tgp_steptanh13_p1', 'power_tgp_sal2_p1', 'boston_svgp_p1'):
    g = Golden(name)
    model = build_from_golden(g, dev)
    X, Y = g.t('X'), g.t('Y')
    sl = D.local_slice(X.shape[0], rank, world)
    model.global_batch_rows = X.shape[0]
    ELBO, ELL, KLD = model.ELBO(X[sl].to(dev), Y[sl].to(dev))
    (-ELBO).backward()
    e = rel_err(ELBO.detach().cpu(), g.t('ELBO'))
    gerr = {}
    for n, prm in model.named_parameters():
        ref = -g.t('grad:' + n)
        if float(ref.norm()) > 0:
            gerr[n] = rel_err(prm.grad.detach().cpu().reshape(ref.shape), ref)
    worst = max(worst, e, max(gerr.values()))
    if rank == 0:
        print('%-28s world %d  ELBO rel %.2e  worst grad rel %.2e (%s)' % (name, world, e, max(gerr.values()), max(gerr, key=gerr.get)))
ok = worst < 1e-8
if rank == 0:
    print('DIST_PARITY', 'OK' if ok else 'FAIL', 'worst %.2e' % worst)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
