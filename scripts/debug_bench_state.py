import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import tgp_oracle as O
from tests.gpu_util import engine_inputs, make_engine
gen = torch.Generator().manual_seed(bench.SEED)
X, Y = bench.synth(200000, bench.D, gen)
p = bench.param_state(X, gen)
Kzz = O.rbf_ard(p['Z'], p['Z'], p['raw_lengthscale'], p['raw_outputscale'])
ev = torch.linalg.eigvalsh(Kzz)
print('Kzz eig min %.3e max %.3e cond %.3e' % (float(ev[0]), float(ev[-1]), float(ev[-1] / ev[0])))
Lref, info = torch.linalg.cholesky_ex(Kzz)
print('LAPACK info', int(info))
for compute in ('f64', 'tf32x3'):
    eng, theta, _, _ = make_engine(p, 'gauss_nonlinear', 100, 'cuda:0', compute=compute)
    ei = engine_inputs(p, 'cuda:0')
    eng.set_params(ei['Z'], ei['raw_ls'], ei['raw_os'], ei['m'], ei['L_raw'], ei['log_var_noise'], theta)
    kl, status = eng.prepare(0.0)
    print(compute, 'status', int(status.item()), 'kl', float(kl))
    L, Linv, C = eng.export_step()
    print('  L rel err vs LAPACK %.3e' % float((L.cpu() - Lref).norm() / Lref.norm()), ' |Linv L - I| max %.3e' % float((Linv.cpu() @ Lref - torch.eye(1024, dtype=torch.float64)).abs().max()))
    xb, yb = X[:65536].cuda(), Y[:65536].view(-1).cuda()
    mu, v = eng.qf_forward(xb)
    print('  mu norm %.10e v min %.6e v sum %.10e' % (float(mu.norm()), float(v.min()), float(v.sum())))
