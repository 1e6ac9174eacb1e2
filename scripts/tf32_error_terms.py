"""Where does the 3xTF32 mode lose its accuracy at cfg4 (ELBO 2.9e-5 from FP64)?  The marginals (mu, v) of
the same rows are computed in 'f64' and in 'tf32x3'; the expected log-likelihood is then evaluated in FP64 from each mixture
(mu_tf, v_64), (mu_64, v_tf), (mu_tf, v_tf), which splits the ELBO error into its mean and variance parts; the variance error is
set against the cancellation in v = s - |a|^2 + |b|^2.
  python scripts/tf32_error_terms.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tests.gpu_util import engine_inputs, make_engine

dev = 'cuda:0'
WL = bench.WORKLOADS['cfg4']
gen = torch.Generator().manual_seed(bench.SEED)
X, Y = bench.synth(WL, 200000, gen)
p = bench.param_state(WL, X, gen)
xb, yb = X[:65536].to(dev), Y[:65536].view(-1).to(dev)
scale = WL['N'] / 65536
out = {}
for mode in ('f64', 'tf32x3', 'i8crt'):
    eng, theta, _, _ = make_engine(p, WL['likelihood'], 100, dev, compute=mode)
    ei = engine_inputs(p, dev)
    eng.set_params(ei['Z'], ei['raw_ls'], ei['raw_os'], ei['m'], ei['L_raw'], ei['log_var_noise'], theta)
    eng.prepare(0.0)
    mu, v = eng.qf_forward(xb)
    out[mode] = (mu.clone(), v.clone(), eng, theta)
mu64, v64, eng64, _ = out['f64']


def ell(mu, v):
    rb = eng64.new_reduce_buffer()
    rows, _, _, _ = eng64.ell_forward(mu.contiguous(), v.contiguous(), yb, None, scale, rb, want_grad=False)
    return float(rows.sum()) * scale


ref = ell(mu64, v64)
s = float(torch.nn.functional.softplus(torch.as_tensor(p['raw_outputscale'])))
print('cfg4, 65536 rows: ELL (f64) = %.10e;  outputscale s = %.3f, median v = %.3e  (v = s - |a|^2 + |b|^2: cancellation factor s / v ~ %.0f)'
      % (ref, s, float(v64.median()), s / float(v64.median())))
for mode in ('tf32x3', 'i8crt'):
    mu, v = out[mode][0], out[mode][1]
    dmu = (mu - mu64).abs()
    dv = (v - v64).abs()
    print('%-7s max |dmu| %.2e (rel to max|mu| %.1e)   max |dv| %.2e, median |dv| / v %.2e, max |dv| / s %.2e'
          % (mode, float(dmu.max()), float(dmu.max() / mu64.abs().max()), float(dv.max()), float((dv / v64.abs()).median()), float(dv.max()) / s))
    for label, a, b in (('mu from %s, v from f64' % mode, mu, v64), ('mu from f64, v from %s' % mode, mu64, v), ('both from %s' % mode, mu, v)):
        e = ell(a, b)
        print('    ELL with %-28s rel. diff %.2e' % (label, abs(e - ref) / abs(ref)))

# What if mu were evaluated as the reference does, mu = K_xz (L^-T m) (sparse_MF_SP.py:354-355), with the FP32-accurate K the tensor-core
# mode generates and FP64 accumulation, instead of from the FP32 rows a = L^-1 k ?
_, Linv, _ = eng64.export_step()
M = WL['M']
u = Linv.t() @ torch.as_tensor(p['m']).to(dev).double()
from oracle import tgp_oracle as O
K = O.rbf_ard(xb[:8192].cpu(), torch.as_tensor(p['Z']), torch.as_tensor(p['raw_lengthscale']), torch.as_tensor(p['raw_outputscale'])).to(dev)
mu_ref = out['f64'][0][:8192]
for label, Kx in (('FP64 K', K), ('K rounded to FP32', K.float().double())):
    mu_alt = Kx @ u
    print('mu = K (L^-T m) with %-18s max |dmu| %.2e (rel to max|mu| %.1e)' % (label, float((mu_alt - mu_ref).abs().max()),
                                                                                 float((mu_alt - mu_ref).abs().max() / mu_ref.abs().max())))
