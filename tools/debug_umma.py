"""Debug: print what the tcgen05 GEMM core returns for tiny single-tile problems (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from distributed_reinforcement_learning_b200 import _native as N
np.set_printoptions(precision=4, linewidth=200, suppress=True)

def run(core, bn, a_km, b_km, M, Nn, K, splits=1, pattern="rand"):
    rng = np.random.default_rng(0)
    if pattern == "rand":
        A = rng.standard_normal((M, K)).astype(np.float32); B = rng.standard_normal((K, Nn)).astype(np.float32)
    elif pattern == "ones":
        A = np.ones((M, K), np.float32); B = np.ones((K, Nn), np.float32)
    elif pattern == "rowid":   # A[m,k] = m+1 for k == 0 else 0 ; B = identity-like -> C[m,n] = (m+1) * [n == 0]
        A = np.zeros((M, K), np.float32); A[:, 0] = np.arange(M) + 1
        B = np.zeros((K, Nn), np.float32); B[0, :] = np.arange(Nn) + 1
    a_in = np.ascontiguousarray(A if a_km else A.T); b_in = np.ascontiguousarray(B.T if b_km else B)
    out = np.full((splits, M + 1, Nn), -7.0, np.float32)
    N.check(N.lib.drl_debug_gemm(core, bn, a_km, b_km, M, Nn, K, splits, N.ptr(a_in), N.ptr(b_in), N.ptr(out)))
    got = out[:, :M].astype(np.float64).sum(0); ref = A.astype(np.float64) @ B.astype(np.float64)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    print("core %d bn %d a_km %d b_km %d M%d N%d K%d %s: err %.3e zeros %d/%d nan %d" % (
        core, bn, a_km, b_km, M, Nn, K, pattern, err, int((got == 0).sum()), got.size, int(np.isnan(got).sum())))
    if err > 1e-4:
        print("  got[0:4,0:8]\n", got[0:4, 0:8]); print("  ref[0:4,0:8]\n", ref[0:4, 0:8])
        print("  got[64:66,0:8]\n", got[64:66, 0:8]); print("  colsum row:", out[0, M, :8])
    return err

for core in (1, 2):
    run(core, 32, 1, 0, 128, 32, 32, pattern="ones")
run(2, 32, 1, 0, 128, 32, 32, pattern="rowid")
run(2, 32, 1, 0, 128, 32, 32)
run(2, 32, 1, 1, 128, 32, 32)
run(2, 32, 0, 0, 128, 32, 32)
run(2, 32, 0, 1, 128, 32, 32)
run(2, 32, 1, 0, 128, 32, 8, pattern="ones")
run(2, 64, 1, 0, 256, 128, 96)
