"""Prints per-quantity parity residuals of the CUDA path for one golden fixture (GPU box)."""
import sys, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.golden_util import Golden, rel_err
from tests.test_gpu_parity import _run_cuda_elbo
from oracle import tgp_oracle as O
warnings.simplefilter('ignore')
for name in sys.argv[1:]:
    g = Golden(name)
    out = _run_cuda_elbo(g)
    print(name, 'ELBO', rel_err(out['ELBO'].detach().cpu(), g.t('ELBO')), 'ELL', rel_err(out['ELL'].detach().cpu(), g.t('ELL')),
          'mu', rel_err(out['mu'].cpu(), g.t('mu')), 'v', rel_err(out['v'].cpu(), g.t('v')), 'vmin', float(out['v'].min()))
    p = g.oracle_params('train')
    rows = O.elbo(g.t('X'), g.t('Y').view(-1), p, g.meta['N'], g.meta['likelihood'], g.meta['n_quad'])[3]
    d = (out['rows'].cpu() - rows).abs()
    print('  rows rel', rel_err(out['rows'].cpu(), rows), 'max abs', float(d.max()), 'at', int(d.argmax()), float(rows[d.argmax()]))
    for k, gr in g.ref_grads().items():
        print('  grad %-18s %.3e  (norm %.3e)' % (k, rel_err(out['grads'][k].detach().cpu(), gr), float(gr.norm())))
