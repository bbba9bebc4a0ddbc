"""K4: Pippenger MSM (G1, G2) and the fixed-base kernel vs the oracle (bit-exact affine output)."""
import random

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import codec
from oracle import groth16 as og
from tests.util import fr_np, random_g1, random_g2

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[0, 1], ids=["plain", "window_tables"], autouse=True)
def msm_mode(request):
    """Every MSM test runs in both modes: plain per-window buckets, or the window tables (2^(c w) P, one
    shared bucket set)."""
    import fawkes_crypto_b200 as fb
    fb.native.lib.fb_set_msm_tables(request.param)
    yield request.param
    fb.native.lib.fb_set_msm_tables(-1)


def gpu_msm(ctx, group, bases, scalars):
    import fawkes_crypto_b200 as fb
    enc = codec.g1_raw if group == 1 else codec.g2_raw
    braw = np.frombuffer(b"".join(enc(p) for p in bases), dtype=np.uint8).copy()
    sc = fr_np(scalars)
    out = np.zeros(64 if group == 1 else 128, dtype=np.uint8)
    fb.native.check(fb.native.lib.fb_test_msm(ctx.handle, group, braw.ctypes.data, sc.ctypes.data, len(bases),
                                              out.ctypes.data, 1, None))
    return (codec.g1_unraw if group == 1 else codec.g2_unraw)(out.tobytes())


@pytest.mark.parametrize("n", [1, 2, 31, 200, 3000])
def test_msm_g1(ctx, n):
    rng = random.Random(n)
    _, bases = random_g1(rng, n)
    scalars = [rng.randrange(bn.R) for _ in range(n)]
    ref = bn.to_affine(bn.OPS1, og.msm(bn.OPS1, bases, scalars))
    assert gpu_msm(ctx, 1, bases, scalars) == ref


@pytest.mark.parametrize("n", [1, 3, 150, 1200])
def test_msm_g2(ctx, n):
    rng = random.Random(1000 + n)
    _, bases = random_g2(rng, n)
    scalars = [rng.randrange(bn.R) for _ in range(n)]
    ref = bn.to_affine(bn.OPS2, og.msm(bn.OPS2, bases, scalars))
    assert gpu_msm(ctx, 2, bases, scalars) == ref


def test_msm_edge_cases(ctx):
    """zeros, ones, r-1, repeated bases (forces the doubling branch), opposite points
    (forces the infinity branch), points at infinity among the bases."""
    rng = random.Random(77)
    _, pts = random_g1(rng, 40)
    g = pts[0]
    bases = pts + [g, g, g, bn.pt_neg(bn.OPS1, g), None, None, g]
    scalars = ([0, 1, bn.R - 1, 2, 1, 1, 0] + [rng.randrange(bn.R) for _ in range(33)] +
               [5, 5, 5, 5, 12345, 0, 1 << 253])
    assert len(bases) == len(scalars)
    ref = bn.to_affine(bn.OPS1, og.msm(bn.OPS1, bases, scalars))
    assert gpu_msm(ctx, 1, bases, scalars) == ref
    # all-zero scalars -> infinity (all-zero raw encoding)
    assert gpu_msm(ctx, 1, pts, [0] * len(pts)) is None
    # witness-like skew: almost everything 0/1
    sk = [rng.choice([0, 1, 1, 1, 2]) for _ in range(len(pts))]
    assert gpu_msm(ctx, 1, pts, sk) == bn.to_affine(bn.OPS1, og.msm(bn.OPS1, pts, sk))


def test_msm_linearity_large(ctx):
    """2^18 points: MSM(s) + MSM(t) == MSM(s+t) and MSM with scalars k_i^-1-free identity:
    bases k_i*G from the fixed-base kernel, so MSM(s) == (sum s_i k_i) * G exactly."""
    import fawkes_crypto_b200 as fb
    n = 1 << 18
    rng = np.random.default_rng(11)
    k = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    k[:, 3] &= np.uint64((1 << 60) - 1)
    s = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    s[:, 3] &= np.uint64((1 << 60) - 1)
    bases = np.zeros((n, 64), dtype=np.uint8)
    fb.native.check(fb.native.lib.fb_test_fixed_base(ctx.handle, 1, k.ctypes.data, n, bases.ctypes.data))
    out = np.zeros(64, dtype=np.uint8)
    fb.native.check(fb.native.lib.fb_test_msm(ctx.handle, 1, bases.ctypes.data, s.ctypes.data, n, out.ctypes.data, 1, None))
    kb, sb = k.tobytes(), s.tobytes()
    tot = 0
    for i in range(n):
        tot += codec.fr_unraw(kb[32 * i:32 * i + 32]) * codec.fr_unraw(sb[32 * i:32 * i + 32])
    assert codec.g1_unraw(out.tobytes()) == bn.pt_mul(bn.OPS1, bn.G1_GEN, tot % bn.R)


def test_msm_skewed_large(ctx):
    """2^17 points with a real-witness-like scalar distribution: 40 % ones, 10 % twos, 10 % zeros, a few r-1, the rest
    random.  One bucket then holds tens of thousands of entries: equal-length tasks spread it over hundreds of threads and
    the heavy-bucket path (k_bucket_heavy: a CTA per bucket, strided sums + tree) folds its partial sums.  Checked with
    the scalar identity MSM(s) == (sum s_i k_i) * G on bases k_i * G."""
    import fawkes_crypto_b200 as fb
    n = 1 << 17
    rng = np.random.default_rng(23)
    k = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    k[:, 3] &= np.uint64((1 << 60) - 1)
    s = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    s[:, 3] &= np.uint64((1 << 60) - 1)
    kind = rng.random(n)
    one, two, rm1 = fr_np([1])[0], fr_np([2])[0], fr_np([bn.R - 1])[0]
    s[kind < 0.4] = one
    s[(kind >= 0.4) & (kind < 0.5)] = two
    s[(kind >= 0.5) & (kind < 0.6)] = 0
    s[(kind >= 0.6) & (kind < 0.601)] = rm1
    bases = np.zeros((n, 64), dtype=np.uint8)
    fb.native.check(fb.native.lib.fb_test_fixed_base(ctx.handle, 1, k.ctypes.data, n, bases.ctypes.data))
    out = np.zeros(64, dtype=np.uint8)
    fb.native.check(fb.native.lib.fb_test_msm(ctx.handle, 1, bases.ctypes.data, s.ctypes.data, n, out.ctypes.data, 1, None))
    kb, sb = k.tobytes(), s.tobytes()
    tot = 0
    for i in range(n):
        tot += codec.fr_unraw(kb[32 * i:32 * i + 32]) * codec.fr_unraw(sb[32 * i:32 * i + 32])
    assert codec.g1_unraw(out.tobytes()) == bn.pt_mul(bn.OPS1, bn.G1_GEN, tot % bn.R)


@pytest.mark.parametrize("group", [1, 2])
def test_fixed_base(ctx, group):
    import fawkes_crypto_b200 as fb
    rng = random.Random(5 + group)
    ks = [0, 1, 2, bn.R - 1, 255, 256, 1 << 248] + [rng.randrange(bn.R) for _ in range(40)]
    sc = fr_np(ks)
    psz = 64 if group == 1 else 128
    out = np.zeros((len(ks), psz), dtype=np.uint8)
    fb.native.check(fb.native.lib.fb_test_fixed_base(ctx.handle, group, sc.ctypes.data, len(ks), out.ctypes.data))
    o, gen, dec = (bn.OPS1, bn.G1_GEN, codec.g1_unraw) if group == 1 else (bn.OPS2, bn.G2_GEN, codec.g2_unraw)
    for i, kk in enumerate(ks):
        assert dec(out[i].tobytes()) == bn.pt_mul(o, gen, kk), i
