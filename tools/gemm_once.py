"""One 8192^3 fp64 GEMM on the library's DMMA kernel (the trailing-update kernel of the Cholesky), for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gp2d_b200 as gp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
A = torch.randn(n, n, dtype=torch.float64, device="cuda")
B = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2):
    C = gp.matmul(A, B)
torch.cuda.synchronize()
print("checksum", float(C[0, 0]))
