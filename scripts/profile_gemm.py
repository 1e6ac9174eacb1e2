"""Launches the FP64 batch-contraction GEMM shapes of one 32768-row chunk of the cfg4 step (for ncu captures):
forward A = K Linv^T (lower-triangular B operand), forward B = A L_S, backward-weight Gbar += Abar^T K (split-K), and —
as the yardstick the roofline is quoted against — one cuBLAS DGEMM of the same dense size through torch.mm."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tgp.pytorch_b200.engine import debug_gemm
dev = 'cuda:0'
R, M = 32768, 1024
f64 = torch.float64
K = torch.rand(R, M, dtype=f64, device=dev)
Linv = torch.randn(M, M, dtype=f64, device=dev).tril()
LS = torch.randn(M, M, dtype=f64, device=dev).tril()
AB = torch.zeros(R, 2 * M, dtype=f64, device=dev)
Abar = torch.randn(R, M, dtype=f64, device=dev)
G = torch.zeros(M, M, dtype=f64, device=dev)
Dn = torch.randn(M, M, dtype=f64, device=dev)
for it in range(2):
    debug_gemm(K, Linv, AB, R, M, M, M, M, 2 * M, 0, 0, b_tri=1)                       # A = K Linv^T
    debug_gemm(AB, LS, AB[:, M:], R, M, M, 2 * M, M, 2 * M, 0, 1, b_tri=2)               # B = A L_S
    debug_gemm(Abar, K, G, M, M, R, M, M, M, 1, 1, beta=1.0, c_lower=1)                  # Gbar += tril(Abar^T K), unsplit
    out = torch.mm(K, Dn)                                                                # cuBLAS, dense 32768 x 1024 x 1024
torch.cuda.synchronize()
print('ok')
