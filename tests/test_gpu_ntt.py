"""K3: evaluation-domain transforms and the H pipeline vs the oracle (bit-exact)."""
import random

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import groth16 as og
from tests.util import fr_np, fr_list

pytestmark = pytest.mark.gpu


def gpu_ntt(ctx, log_n, kind, vals):
    import fawkes_crypto_b200 as fb
    a = fr_np(vals)
    fb.native.check(fb.native.lib.fb_test_ntt(ctx.handle, log_n, kind, a.ctypes.data))
    return fr_list(a)


@pytest.mark.parametrize("log_n", [1, 2, 3, 5, 8, 11, 12, 13, 14])
def test_transforms_match_oracle(ctx, log_n):
    rng = random.Random(log_n)
    n = 1 << log_n
    vals = [rng.randrange(bn.R) for _ in range(n)]
    vals[0] = 0
    vals[-1] = bn.R - 1
    m, exp, omega = og.domain_params(n)
    for kind, fn in ((0, og.fft), (1, og.ifft), (2, og.coset_fft), (3, og.icoset_fft)):
        ref = list(vals)
        fn(ref, omega, exp)
        assert gpu_ntt(ctx, log_n, kind, vals) == ref, f"kind {kind} log_n {log_n}"


@pytest.mark.parametrize("log_n", [1, 4, 10, 12, 13, 15])
def test_h_pipeline_matches_oracle(ctx, log_n):
    import fawkes_crypto_b200 as fb
    rng = random.Random(100 + log_n)
    n = 1 << log_n
    a = [rng.randrange(bn.R) for _ in range(n)]
    b = [rng.randrange(bn.R) for _ in range(n)]
    c = [rng.randrange(bn.R) for _ in range(n)]
    ref = og.h_coefficients(a, b, c)
    an, bnp, cn = fr_np(a), fr_np(b), fr_np(c)
    out = np.zeros((n - 1, 4), dtype=np.uint64)
    fb.native.check(fb.native.lib.fb_test_h(ctx.handle, log_n, an.ctypes.data, bnp.ctypes.data, cn.ctypes.data,
                                            out.ctypes.data, None))
    assert fr_list(out) == ref


def test_roundtrip_large(ctx):
    """2^20: icoset_fft(coset_fft(x)) == x and ifft(fft(x)) == x (size-independent property)."""
    import fawkes_crypto_b200 as fb
    log_n = 20
    rng = np.random.default_rng(5)
    x = rng.integers(0, 1 << 62, size=(1 << log_n, 4), dtype=np.uint64)
    x[:, 3] &= np.uint64((1 << 60) - 1)
    for fwd, inv in ((0, 1), (2, 3)):
        y = x.copy()
        fb.native.check(fb.native.lib.fb_test_ntt(ctx.handle, log_n, fwd, y.ctypes.data))
        assert not np.array_equal(x, y)
        fb.native.check(fb.native.lib.fb_test_ntt(ctx.handle, log_n, inv, y.ctypes.data))
        assert np.array_equal(x, y)


@pytest.mark.parametrize("log_n,g", [(16, 1), (17, 2), (18, 3), (16, 3)])
def test_distributed_h_pipeline_layouts(ctx, log_n, g):
    """Four-step (cyclic/block) H pipeline with 2^g virtual ranks on one GPU == the single-GPU pipeline
    (itself checked against the oracle above).  The NCCL transport is exercised by bench.py --gpus N."""
    import fawkes_crypto_b200 as fb
    rng = np.random.default_rng(log_n * 10 + g)
    n = 1 << log_n
    arrs = []
    for _ in range(3):
        x = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
        x[:, 3] &= np.uint64((1 << 60) - 1)
        arrs.append(x)
    ref = np.zeros((n - 1, 4), dtype=np.uint64)
    fb.native.check(fb.native.lib.fb_test_h(ctx.handle, log_n, arrs[0].ctypes.data, arrs[1].ctypes.data,
                                            arrs[2].ctypes.data, ref.ctypes.data, None))
    out = np.zeros((n - 1, 4), dtype=np.uint64)
    fb.native.check(fb.native.lib.fb_test_dist_h(ctx.handle, log_n, g, arrs[0].ctypes.data, arrs[1].ctypes.data,
                                                 arrs[2].ctypes.data, out.ctypes.data))
    assert np.array_equal(out, ref)
