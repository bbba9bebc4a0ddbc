"""Gate-blob ingest on the GPU (fb_circuit_from_raw_gates_gpu, csrc/ingest.cu) against the host parser
(fb_circuit_from_raw_gates, csrc/pk.cu) -- same CSR, same coefficient dictionary, same error behaviour as the
reference's stream (circuit/r1cs/cs.rs:184-223: a gate exists only if all three LCs deserialize; a coefficient
>= r is "Wrong raw integer", ff-uint_derive/src/lib.rs:687-702)."""
import ctypes as C

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import codec
from oracle import synth
from tests.util import random_gate_blob

pytestmark = pytest.mark.gpu


def raw_csr(fb, circ):
    """The circuit's CSR exactly as stored: per matrix (rowptr, col, cidx, coef table)."""
    lib, out, sh = fb.native.lib, [], circ.shape()
    for m in range(3):
        prp, pcl, pci, pct = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        nnz, ncoef = C.c_uint64(), C.c_uint64()
        fb.native.check(lib.fb_circuit_csr(circ.handle, m, C.byref(prp), C.byref(pcl), C.byref(pci), C.byref(nnz),
                                           C.byref(pct), C.byref(ncoef)))
        u32 = lambda p, n: np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(n,)).copy() if n else np.zeros(0, np.uint32)
        tab = (np.ctypeslib.as_array(C.cast(pct, C.POINTER(C.c_uint64)), shape=(ncoef.value, 4)).copy()
               if ncoef.value else np.zeros((0, 4), np.uint64))
        out.append((u32(prp, sh["n_gates"] + 1), u32(pcl, nnz.value), u32(pci, nnz.value), tab))
    return out


def assert_same_circuit(fb, a, b, exact=True):
    assert a.shape() == b.shape()
    for (rp1, cl1, ci1, t1), (rp2, cl2, ci2, t2) in zip(raw_csr(fb, a), raw_csr(fb, b)):
        assert np.array_equal(rp1, rp2) and np.array_equal(cl1, cl2)
        if exact:
            assert np.array_equal(ci1, ci2) and np.array_equal(t1, t2)
        else:       # same coefficient per term, dictionary numbering free
            small = (ci1 < 2) | (ci2 < 2)
            assert np.array_equal(ci1[small], ci2[small])
            assert np.array_equal(t1[ci1[~small] - 2], t2[ci2[~small] - 2])


@pytest.mark.parametrize("n_gates", [1, 2, 37, 5000])
def test_gpu_ingest_equals_host_parser_random_streams(ctx, n_gates):
    import fawkes_crypto_b200 as fb
    raw = random_gate_blob(n_gates, 3, 1000, seed=n_gates, terms=(3, 2, 1) if n_gates % 2 else (5, 0, 4))
    host = fb.Circuit.from_raw_gates(raw, n_gates, 3, 1000)
    dev = fb.Circuit.from_raw_gates(raw, n_gates, 3, 1000, ctx=ctx)
    assert_same_circuit(fb, host, dev)
    assert set(dev.ingest_ms) == {"frame", "upload", "kernels", "download", "brotli", "alloc"}


def test_gpu_ingest_cfg1_merkle_circuit_and_brotli(ctx):
    """The real cfg-1 gate stream (LCs of up to ~55 terms, a few hundred distinct MDS-derived coefficients
    repeated thousands of times) through the brotli entry point."""
    import random
    import fawkes_crypto_b200 as fb
    from oracle import frontend as fe
    rng = random.Random(7)
    gates, inp, aux = fe.merkle_circuit(rng.randrange(bn.R), [rng.randrange(bn.R) for _ in range(32)],
                                        [rng.random() < 0.5 for _ in range(32)])
    raw = b"".join(codec.gate_borsh(g) for g in gates)
    blob = codec.brotli_compress(raw)
    host = fb.Circuit.from_gates_blob(blob, len(gates), 2, len(aux))
    dev = fb.Circuit.from_gates_blob(blob, len(gates), 2, len(aux), ctx=ctx)
    assert_same_circuit(fb, host, dev)
    tab = raw_csr(fb, dev)[0][3]
    assert 0 < len(tab) < 5000 < host.shape()["nnz"]           # the dictionary really dedups


def test_gpu_ingest_large_stream_beyond_host_dictionary_cap(ctx):
    """2^18 gates, ~370k distinct coefficients and 1.8 M terms: same terms either way."""
    import fawkes_crypto_b200 as fb
    n = 1 << 18
    raw = random_gate_blob(n, 2, n, seed=99)
    host = fb.Circuit.from_raw_gates(raw, n, 2, n)
    dev = fb.Circuit.from_raw_gates(raw, n, 2, n, ctx=ctx)
    assert_same_circuit(fb, host, dev, exact=True)             # below the 2^20 cap: identical numbering too


def test_gpu_ingest_error_behaviour_matches_host(ctx):
    import fawkes_crypto_b200 as fb
    n = 64
    raw = bytearray(random_gate_blob(n, 2, 50, seed=5, terms=(3, 3, 1)))
    per_gate = 3 * 4 + 7 * 37

    def both(buf, num_gates):
        res = []
        for kw in ({}, {"ctx": ctx}):
            try:
                c = fb.Circuit.from_raw_gates(bytes(buf), num_gates, 2, 50, **kw)
                res.append(("ok", c.shape()["n_gates"]))
            except fb.native.FbError as e:
                res.append(("err", str(e)))
        assert res[0] == res[1], res
        return res[0]

    assert both(raw, n) == ("ok", n)
    # truncated in the middle of gate 40: the stream holds 40 gates
    assert both(raw[:40 * per_gate + 50], 40) == ("ok", 40)
    assert both(raw[:40 * per_gate + 50], n)[0] == "err"
    # a coefficient >= r in gate 10 (second term of B) ends the stream there
    bad = bytearray(raw)
    off = 10 * per_gate + 4 + 3 * 37 + 4 + 37
    bad[off:off + 32] = bn.R.to_bytes(32, "little")
    assert both(bad, 10) == ("ok", 10)
    assert both(bad, n)[0] == "err"
    # an Index tag of 2 does the same
    bad = bytearray(raw)
    bad[20 * per_gate + 4 + 32] = 2
    assert both(bad, 20) == ("ok", 20)
    # a variable index out of range is a format error that names the gate
    bad = bytearray(raw)
    off = 33 * per_gate + 4 + 37
    bad[off + 32] = 1
    bad[off + 33:off + 37] = (50).to_bytes(4, "little")
    r = both(bad, n)
    assert r[0] == "err" and "gate 33" in r[1] and "aux variable 50" in r[1]
    # empty stream
    assert both(b"", 0) == ("ok", 0)


def test_pk_load_from_blob_uses_gpu_ingest_and_proves(ctx):
    """fb_pk_load (Parameters bytes + brotli blob, the reference's own inputs) goes through the GPU ingest;
    the proof equals the oracle's."""
    import fawkes_crypto_b200 as fb
    from oracle import groth16 as og
    from tests.util import fr_np
    seed = synth.SEED_BASE + 4242
    gates, inp, aux = synth.synth_circuit(300, seed)
    td, r, s = synth.synth_trapdoor(seed)
    P = og.setup(gates, 2, len(aux), td)
    blob = codec.brotli_compress(b"".join(codec.gate_borsh(g) for g in gates))
    pb = codec.bellman_params_bytes(P)
    pk = C.c_void_p()
    fb.native.check(fb.native.lib.fb_pk_load(ctx.handle, fb.native.ptr(pb), len(pb), fb.native.ptr(blob), len(blob),
                                             len(gates), 1, C.byref(pk)))
    wi, wa = fr_np(inp), fr_np(aux)
    out = np.zeros(256, dtype=np.uint8)
    ra, sa = fb.groth16.fr_raw(r), fb.groth16.fr_raw(s)
    fb.native.check(fb.native.lib.fb_prove(ctx.handle, pk, fb.native.ptr(wi), 2, fb.native.ptr(wa), len(aux),
                                           fb.native.ptr(ra), fb.native.ptr(sa), fb.native.ptr(out), None))
    fb.native.lib.fb_pk_free(pk)
    assert out.tobytes() == codec.proof_raw(og.prove(P, gates, inp, aux, r, s))
