"""Kernel time of the CSR x dense path on the config #3 R-MAT matrix (n = 2048)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sparta_b200
from sparta_b200 import synth
scale, n = 16, 2048
N = 1 << scale
r, c = synth.rmat_edges(scale, int(1e-3 * N * N), seed=1)
r, c = synth.pin_shape(r, c, N, N)
rowptr, colind, _ = synth.csr_from_edges(r, c, N)
nnz = len(colind)
B = synth.seeded_B(N, n, seed=2).T.copy()   # row-major N x n
for prec in ("bf16", "tf32"):
    h = sparta_b200.Handle.from_csr(N, N, rowptr, colind, None, precision=prec)
    h.set_B(B, n, n)
    for _ in range(3): h.run()
    ts = [h.run() for _ in range(10)]
    ms = float(np.median(ts))
    es = 2 if prec == "bf16" else 4
    gather = nnz * n * es
    print(f"csr {prec}: {ms:.3f} ms  nnz={nnz}  {2*nnz*n/ms/1e9:.1f} TFLOP/s-eff  L2->SM gather {gather/1e9:.1f} GB = {gather/ms/1e9:.2f} TB/s  "
          f"HBM min {(N*n*es + N*n*4 + nnz*8)/1e9:.2f} GB")
    h.close()
