"""compute-sanitizer target (race / memory checking): one pass through every kernel family at
small sizes -- field self-test, NTT (strided + contiguous tiles), the H pipeline incl. the virtual-rank distributed
layout, G1/G2 MSMs with and without window tables, GPU gate ingest, setup
(full and sharded), a plain prove, a CUDA-graph prove, a batched prove (buckets keyed by proof) and verify.
    compute-sanitizer --tool memcheck  python tools/sanitize_target.py
    compute-sanitizer --tool racecheck python tools/sanitize_target.py      (shared-memory hazards: NTT tiles,
                                       the k_segment_bits / k_bucket_heavy / k_fold_parts trees, scans)
Every result is also compared with the CPU oracle, so a silent corruption would not pass either."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fawkes_crypto_b200 as fb  # noqa: E402
from oracle import cpu  # noqa: E402

lib = fb.native.lib
ctx = fb.Context(0)
rng = np.random.default_rng(7)


def rand_fr(n):
    x = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    x[:, 3] &= np.uint64((1 << 60) - 1)
    return x


# field ops
a, b = rand_fr(257), rand_fr(257)
out = np.zeros_like(a)
for field in (0, 1):
    for op in (0, 1, 2, 3):
        fb.native.check(lib.fb_test_field(ctx.handle, field, op, a.ctypes.data, b.ctypes.data, out.ctypes.data, 257))
# transforms: 2^12 (one strided pass + the contiguous tile pass) and round trips
for lg in (3, 12):
    x = rand_fr(1 << lg)
    y = x.copy()
    for kind in (0, 1, 2, 3):
        fb.native.check(lib.fb_test_ntt(ctx.handle, lg, kind, y.ctypes.data))
    fb.native.check(lib.fb_test_ntt(ctx.handle, lg, 0, y.ctypes.data))
    fb.native.check(lib.fb_test_ntt(ctx.handle, lg, 1, y.ctypes.data))
# H pipeline, single and virtual-rank distributed (2^14 over 4 ranks)
lg = 14
ea, eb, ec = rand_fr(1 << lg), rand_fr(1 << lg), rand_fr(1 << lg)
h1 = np.zeros(((1 << lg) - 1, 4), dtype=np.uint64)
h2 = np.zeros_like(h1)
fb.native.check(lib.fb_test_h(ctx.handle, lg, ea.ctypes.data, eb.ctypes.data, ec.ctypes.data, h1.ctypes.data, None))
fb.native.check(lib.fb_test_dist_h(ctx.handle, lg, 2, ea.ctypes.data, eb.ctypes.data, ec.ctypes.data, h2.ctypes.data))
assert np.array_equal(h1, h2), "distributed H pipeline differs"
# MSMs
n = 3000
k, s = rand_fr(n), rand_fr(n)
s[:50] = 0
s[50:120, :] = fb.groth16.fr_raw(1)          # skewed digits: many scalars equal to one
for group, psz in ((1, 64), (2, 128)):
    bases = np.zeros((n, psz), dtype=np.uint8)
    fb.native.check(lib.fb_test_fixed_base(ctx.handle, group, k.ctypes.data, n, bases.ctypes.data))
    ref = None
    for tables in (0, 1):
        lib.fb_set_msm_tables(tables)
        res = np.zeros(psz, dtype=np.uint8)
        fb.native.check(lib.fb_test_msm(ctx.handle, group, bases.ctypes.data, s.ctypes.data, n, res.ctypes.data, 1, None))
        ref = res if ref is None else ref
        assert np.array_equal(res, ref), ("msm modes differ", group, tables)
    if group == 1:
        cref, _ = cpu.msm_g1(bases, s, 2)
        assert cref == ref.tobytes(), "G1 MSM differs from the CPU oracle"
lib.fb_set_msm_tables(-1)
# circuit -> (sharded) setup -> prove (stream path, graph path, batched path) -> verify; GPU ingest
seed = 0xFA3CE50000 + 4242
occ = cpu.Circuit.synthetic(700, seed)
td = cpu.synth_trapdoor(seed)
circ = fb.Circuit.synthetic(700, seed)
tdi = [fb.groth16.fr_unraw(x) for x in td]
params = fb.setup(circ, ctx, trapdoor=tdi[:5])
pbuf, _ = cpu.setup(occ, td, 2)
assert bytes(params.bellman_bytes) == pbuf.tobytes(), "GPU setup differs from the CPU oracle"
sh = fb.setup(circ, ctx, trapdoor=tdi[:5], shard=1, nshards=3)
assert len(sh.bellman_bytes) == len(params.bellman_bytes)
wi, wa = circ.witness()
want, _, _ = cpu.prove_circuit(pbuf, occ, td[5], td[6], 2)
for graph in (0, 1):
    lib.fb_set_prove_graph(graph)
    _, proof = fb.groth16.prove_with_rs(params, wi, wa, tdi[5], tdi[6], ctx)
    assert proof.to_raw() == want, ("prove differs from the CPU oracle", graph)
for mode, chunk in (("batched", "2"), ("slots", "0")):
    os.environ["FB_BATCH_MODE"], os.environ["FB_BATCH_P"] = mode, chunk
    res = fb.prove_batch(params, [(wi, wa)] * 5, [tdi[5]] * 5, [tdi[6]] * 5, ctx)
    assert all(p.to_raw() == want for _, p in res), ("batched prove differs", mode)
assert fb.verify(params.get_vk(), proof, wi[1:])
params.unload()
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import random_gate_blob  # noqa: E402
raw = random_gate_blob(500, 3, 60, seed=5)
c1 = fb.Circuit.from_raw_gates(raw, 500, 3, 60, ctx)
c2 = fb.Circuit.from_raw_gates(raw, 500, 3, 60)
assert c1.shape() == c2.shape()
print("sanitize target ok")
