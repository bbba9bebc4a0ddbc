"""Run one G1 (and optionally G2) MSM of 2^LOG points -- target for ncu launch lists."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fawkes_crypto_b200 as fb
lib = fb.native.lib
ctx = fb.Context(0)
lg = int(os.environ.get("LOG", "18")); group = int(os.environ.get("GROUP", "1")); reps = int(os.environ.get("REPS", "1"))
n = 1 << lg
rng = np.random.default_rng(1)
def rand_fr(n):
    x = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64); x[:, 3] &= np.uint64((1 << 60) - 1); return x
k, s = rand_fr(n), rand_fr(n)
psz = 64 if group == 1 else 128
bases = np.zeros((n, psz), dtype=np.uint8)
fb.native.check(lib.fb_test_fixed_base(ctx.handle, group, k.ctypes.data, n, bases.ctypes.data))
res = np.zeros(psz, dtype=np.uint8); ms = C.c_float()
fb.native.check(lib.fb_test_msm(ctx.handle, group, bases.ctypes.data, s.ctypes.data, n, res.ctypes.data, reps, C.byref(ms)))
print("msm", lg, group, ms.value, "ms")
