"""Launches the tcgen05 3xTF32 GEMM with the shapes of one 16384-row chunk of the cfg4 step (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float32)
from tgp.pytorch_b200.engine import debug_gemm_tf32x3
dev = 'cuda:0'
R, M = 16384, 1024
K = torch.rand(R, M, device=dev)
W = torch.randn(2 * M, M, device=dev)
AB = torch.randn(R, 2 * M, device=dev)
Wt = torch.randn(M, 2 * M, device=dev)
KT = torch.rand(M, R, device=dev)
PT = torch.randn(M, R, device=dev)
out = torch.zeros(R, 2 * M, device=dev)
kbar = torch.zeros(R, M, device=dev)
G = torch.zeros(M, M, dtype=torch.float64, device=dev)
for it in range(3):
    debug_gemm_tf32x3(K, W, out, tri_mode=1, tri_rows=M)                    # forward  [A|B] = K [Linv;C]^T
    debug_gemm_tf32x3(AB, Wt, kbar, tri_mode=2, tri_rows=M)                  # backward data
    debug_gemm_tf32x3(PT, KT, G, out_mode=1, lower_rows=0, splitk=8)         # backward weight (reduction over rows)
torch.cuda.synchronize()
print('ok')
