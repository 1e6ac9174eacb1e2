"""Multi-GPU parity (run under torchrun): every rank evaluates its row slice of a golden fixture's minibatch through the
class API; the all-reduced ELBO and gradients must equal the reference's single-process values (1e-10).  Covers the
input-dependent flows too (MLP gradients are summed over ranks; the dropout-on fixture slices the recorded masks by rows).
Prints one DIST_PARITY line and writes gpurun_out/dist_parity_<world>gpu.json on rank 0."""
import json, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from tests.golden_util import Golden, rel_err
from tests.model_util import build_from_golden, set_dropout_mode
from tgp.pytorch_b200 import dist as D
from tgp.pytorch_b200.dsp import config as cg

warnings.simplefilter('ignore')
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = 'cuda:%d' % local
dist.init_process_group('nccl', device_id=torch.device(dev))
report, worst = {}, 0.0
for sync in (True, False):
    cg.sync_elbo_in_forward = sync          # True: global ELBO returned by the forward; False: one collective per step
    for name in ('synth_reg_d8_m64_p1', 'boston_tgp_steptanh13_p1', 'power_tgp_sal2_p1', 'boston_svgp_p1', 'synth_reg_d8_m1024_p1',
                 'power_idtgp_nodrop_p1', 'boston_idtgp_drop_p1'):
        g = Golden(name)
        model = build_from_golden(g, dev)
        set_dropout_mode(model, g)
        X, Y = g.t('X'), g.t('Y')
        sl = D.local_slice(X.shape[0], rank, world)
        for layer in model.G_matrix[0].flow_arr if hasattr(model, 'G_matrix') and len(model.G_matrix) else []:
            if getattr(layer, 'dropout_masks', None) is not None:
                layer.dropout_masks = layer.dropout_masks[:, :, sl].contiguous()        # this rank's rows of the recorded masks
        model.global_batch_rows = X.shape[0]
        ELBO, ELL, KLD = model.ELBO(X[sl].to(dev), Y[sl].to(dev))
        (-ELBO).backward()
        val = ELBO.detach() if sync else model.last_global_elbo()
        e = rel_err(val.cpu(), g.t('ELBO'))
        gerr = {}
        for n, prm in model.named_parameters():
            if 'grad:' + n not in g.z.files:
                gerr.update({k: v for k, v in g.grad_errors({'L_raw': -prm.grad.detach()[0]}).items() if k.startswith('L_raw.')})
                continue
            ref = -g.t('grad:' + n)
            if float(ref.norm()) > 0:
                gerr[n] = rel_err(prm.grad.detach().cpu().reshape(ref.shape), ref)
        w = max(e, max(gerr.values()))
        worst = max(worst, w)
        report['%s|sync=%d' % (name, sync)] = {'elbo': e, 'worst_grad': max(gerr.values()), 'worst_grad_name': max(gerr, key=gerr.get)}
        if rank == 0:
            print('%-28s world %d sync %d  ELBO rel %.2e  worst grad rel %.2e (%s)' % (name, world, sync, e, max(gerr.values()), max(gerr, key=gerr.get)))
# multiclass (one GP per class, Monte-Carlo softmax likelihood): the marginals' backward all-reduces one packed buffer per class,
# the flow scalars are summed across ranks; the recorded N(0,1) draws are sliced by rows like the data
from tests.golden_util import MulticlassGolden, multiclass_names
from tests.test_gpu_multiclass import _build as build_multiclass
cg.sync_elbo_in_forward = True
for name in multiclass_names():
    g = MulticlassGolden(name)
    import tests.test_gpu_multiclass as T
    T.DEV = dev
    model = build_multiclass(g)
    X, Y = g.t('X'), torch.tensor(np.asarray(g.z['Y']))
    sl = D.local_slice(X.shape[0], rank, world)
    model.global_batch_rows = X.shape[0]
    model.likelihood.mc_noise = g.t('eps')[:, :, sl].contiguous().to(dev)
    ELBO, ELL, KLD = model.ELBO(X[sl].to(dev), Y[sl].to(dev))
    (-ELBO).backward()
    e = rel_err(ELBO.detach().cpu(), g.t('ELBO'))
    gerr = {n: rel_err(-prm.grad.detach().cpu().reshape(-1), g.t('grad:' + n).reshape(-1)) for n, prm in model.named_parameters()}
    w = max(e, max(gerr.values()))
    worst = max(worst, w)
    report['%s|multiclass' % name] = {'elbo': e, 'worst_grad': max(gerr.values()), 'worst_grad_name': max(gerr, key=gerr.get)}
    if rank == 0:
        print('%-28s world %d multiclass  ELBO rel %.2e  worst grad rel %.2e (%s)' % (name, world, e, max(gerr.values()), max(gerr, key=gerr.get)))
t = torch.tensor([worst], dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
worst = float(t.item())
ok = worst < 1e-10
if rank == 0:
    print('DIST_PARITY', 'OK' if ok else 'FAIL', 'worst %.2e' % worst)
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump({'world': world, 'worst': worst, 'ok': ok, 'fixtures': report}, open('gpurun_out/dist_parity_%dgpu.json' % world, 'w'), indent=1)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
