"""Launches the batch-contraction GEMM shapes of one 8192-row chunk of the cfg4 step (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tgp.pytorch_b200.engine import debug_gemm
dev = 'cuda:0'
R, M = 8192, 1024
K = torch.randn(R, M, dtype=torch.float64, device=dev)
W = torch.randn(M, M, dtype=torch.float64, device=dev)
AB = torch.randn(R, 2 * M, dtype=torch.float64, device=dev)
out = torch.zeros(R, M, dtype=torch.float64, device=dev)
G = torch.zeros(M, M, dtype=torch.float64, device=dev)
for it in range(3):
    debug_gemm(K, W, out, R, M, M, M, M, M, 0, 0)                       # forward B-part: dense NT
    debug_gemm(AB, W, out, R, M, M, 2 * M, M, M, 0, 1)                  # backward data: NN
    debug_gemm(AB, K, G, M, M, R, 2 * M, M, M, 1, 1, beta=1.0)          # backward weight: TN, reduction over rows
torch.cuda.synchronize()
print('ok')
