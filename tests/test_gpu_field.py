"""K1: Fr/Fq device arithmetic vs Python integers, through fb_test_field (bit-exact)."""
import random

import numpy as np
import pytest

from oracle import bn254 as bn
from tests.util import fr_np, fq_np, fr_list, fq_list, edge_values

pytestmark = pytest.mark.gpu

FIELDS = {0: (bn.R, fr_np, fr_list), 1: (bn.P, fq_np, fq_list)}


def run_op(ctx, field, op, a, b):
    import fawkes_crypto_b200 as fb
    out = np.zeros_like(a)
    fb.native.check(fb.native.lib.fb_test_field(ctx.handle, field, op, a.ctypes.data,
                                                b.ctypes.data if b is not None else None, out.ctypes.data, a.shape[0]))
    return out


@pytest.mark.parametrize("field", [0, 1])
def test_field_ops_match_integers(ctx, field):
    mod, to_np, to_list = FIELDS[field]
    rng = random.Random(1234 + field)
    n = 5000
    av = edge_values(mod, rng, n)
    bv = edge_values(mod, rng, n)
    rng.shuffle(bv)
    # make sure every edge pairs with every edge at least for the first 14x14 block
    av[:196] = [x for x in av[:14] for _ in range(14)]
    bv[:196] = [y for _ in range(14) for y in edge_values(mod, rng, 14)]
    a, b = to_np(av), to_np(bv)
    assert to_list(run_op(ctx, field, 0, a, b)) == [x * y % mod for x, y in zip(av, bv)]      # PTX mul
    assert to_list(run_op(ctx, field, 4, a, b)) == [x * y % mod for x, y in zip(av, bv)]      # portable mul
    assert to_list(run_op(ctx, field, 1, a, b)) == [(x + y) % mod for x, y in zip(av, bv)]
    assert to_list(run_op(ctx, field, 2, a, b)) == [(x - y) % mod for x, y in zip(av, bv)]
    # dedicated square and the two-products-one-reduction form used by the mixed add
    assert to_list(run_op(ctx, field, 5, a, b)) == [x * x % mod for x in av]
    assert to_list(run_op(ctx, field, 6, a, b)) == [(x * y - y * x * x) % mod for x, y in zip(av, bv)]
    if field == 1:  # Fq2 product with lazy reduction: (x + y u)(y + x^2 u), u^2 = -1, reported as c0 - c1
        want = [((x * y - y * x * x) - (x * x * x + y * y)) % mod for x, y in zip(av, bv)]
        assert to_list(run_op(ctx, field, 7, a, b)) == want
        # Fq2 a*b - c*d with two shared reductions equals the four-product form (difference is zero)
        assert not np.any(run_op(ctx, field, 8, a, b))


@pytest.mark.parametrize("field", [0, 1])
def test_field_inverse(ctx, field):
    mod, to_np, to_list = FIELDS[field]
    rng = random.Random(99)
    av = edge_values(mod, rng, 300)
    got = to_list(run_op(ctx, field, 3, to_np(av), None))
    assert got == [pow(x, -1, mod) if x else 0 for x in av]   # inv(0) -> 0 (reference: None)


def test_ptx_mul_equals_portable_mul_on_raw_limbs(ctx):
    """Same unreduced-but-valid limb patterns through both multipliers (1e5 samples)."""
    rng = np.random.default_rng(7)
    n = 100000
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * 2 + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    b = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * 2 + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)   # keep operands < p (top limb of p is 0x3064...)
    b[:, 3] &= np.uint64((1 << 60) - 1)
    for field in (0, 1):
        assert np.array_equal(run_op(ctx, field, 0, a, b), run_op(ctx, field, 4, a, b))
