"""C-ABI surface and host-side logic, no GPU needed: the library loads, exports every symbol
include/fawkes_b200.h declares, and its host code (gate parsing, synthetic generator, framing,
pairing verifier, partial-sum combine) agrees with the oracle."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import codec
from oracle import groth16 as og
from oracle import synth
from tests.util import fr_np, fr_list

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import fawkes_crypto_b200 as fb
    hdr = open(os.path.join(ROOT, "include", "fawkes_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    for n in sorted(names):
        assert hasattr(fb.native.lib, n), f"{n} declared in the header but not exported"
    assert names == set(fb.native.SIGNATURES), names ^ set(fb.native.SIGNATURES)


def test_no_cpu_fallback_without_device():
    import fawkes_crypto_b200 as fb
    if fb.native.lib.fb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(fb.native.FbError) as e:
        fb.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "fawkes-crypto_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "oracle/" not in src.replace("oracle/synth.py is the Python restatement", ""), f


def test_synthetic_generator_matches_python_restatement():
    import fawkes_crypto_b200 as fb
    for n_rows in (3, 10, 257):
        seed = synth.SEED_BASE + 50 + n_rows
        gates, inp, aux = synth.synth_circuit(n_rows, seed)
        c = fb.Circuit.synthetic(n_rows, seed)
        wi, wa = c.witness()
        assert fr_list(wi) == inp and fr_list(wa) == aux
        sh = c.shape()
        assert sh["n_gates"] == len(gates) and sh["nnz"] == sum(len(x) for g in gates for x in g)
        import bench
        rp, cl, cf = bench.expand_csr(fb, c)
        for m in range(3):
            terms = [(codec.fr_unraw(cf[m][i].tobytes()), int(cl[m][i])) for i in range(len(cl[m]))]
            want = [(cv, idx if tag == 0 else 2 + idx) for g in gates for cv, (tag, idx) in g[m]]
            assert terms == want
        td = np.zeros((7, 4), dtype=np.uint64)
        fb.native.check(fb.native.lib.fb_synth_trapdoor(seed, td.ctypes.data))
        t, r, s = synth.synth_trapdoor(seed)
        assert fr_list(td) == [t.alpha, t.beta, t.gamma, t.delta, t.tau, r, s]


def test_gate_stream_parsing_and_errors():
    import fawkes_crypto_b200 as fb
    gates, inp, aux = synth.synth_circuit(30, 7)
    raw = b"".join(codec.gate_borsh(g) for g in gates)
    blob = codec.brotli_compress(raw)
    c = fb.Circuit.from_gates_blob(blob, len(gates), 2, len(aux))
    assert c.shape() == {"n_in": 2, "n_aux": len(aux), "n_gates": len(gates), "nnz": sum(len(x) for g in gates for x in g)}
    # truncated stream -> fewer gates than announced -> FB_ERR_FORMAT
    with pytest.raises(fb.native.FbError) as e:
        fb.Circuit.from_raw_gates(raw[:-3], len(gates), 2, len(aux))
    assert e.value.code == -3
    # coefficient >= r is "Wrong raw integer": the stream ends there (cs.rs:215-223)
    bad = bytearray(raw)
    bad[4:36] = (bn.R + 5).to_bytes(32, "little")
    with pytest.raises(fb.native.FbError):
        fb.Circuit.from_raw_gates(bytes(bad), len(gates), 2, len(aux))
    # variable index out of range
    with pytest.raises(fb.native.FbError):
        fb.Circuit.from_raw_gates(raw, len(gates), 2, len(aux) - 5)
    # empty circuit
    c0 = fb.Circuit.from_raw_gates(b"", 0, 1, 0)
    assert c0.shape()["n_gates"] == 0


def test_host_parser_equals_python_parse_on_random_streams():
    """The host parser is the yardstick of the GPU ingest (tests/test_gpu_ingest.py): check it term by term
    against the Python restatement of the stream format (oracle codec, cs.rs:184-223) on random streams
    with repeated, +1, -1 and fresh coefficients and ragged LC lengths."""
    import bench
    import fawkes_crypto_b200 as fb
    from tests.util import random_gate_blob
    for n_gates, terms in ((1, (3, 2, 1)), (64, (5, 0, 4)), (300, (3, 3, 1))):
        raw = random_gate_blob(n_gates, 3, 40, seed=n_gates, terms=terms)
        ref = codec.parse_gates(raw)
        assert len(ref) == n_gates
        c = fb.Circuit.from_raw_gates(raw, n_gates, 3, 40)
        rp, cl, cf = bench.expand_csr(fb, c)
        for m in range(3):
            want_cols = [(idx if tag == 0 else 3 + idx) for g in ref for _, (tag, idx) in g[m]]
            want_coef = [cc for g in ref for cc, _ in g[m]]
            assert list(rp[m]) == [0] + list(np.cumsum([len(g[m]) for g in ref]))
            assert list(cl[m]) == want_cols
            assert fr_list(cf[m]) == want_coef


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "synth_rows40.json")))


def test_host_verify_and_framing_on_golden(golden):
    import fawkes_crypto_b200 as fb
    pb = bytes.fromhex(golden["bellman_params_hex"])
    raw = bytes.fromhex(golden["gates_raw_hex"])
    params = fb.Parameters(pb, 38, codec.brotli_compress(raw), [True, False])
    p2 = fb.Parameters.read(params.write())
    assert p2.bellman_bytes == pb and p2.num_gates == 38 and p2.const_tracker == [True, False]
    with pytest.raises(IOError):
        fb.Parameters.read(params.write()[:20])
    proof = fb.Proof.from_raw(bytes.fromhex(golden["proof_raw_hex"]))
    assert proof.serialize().hex() == golden["proof_borsh_hex"]
    assert fb.Proof.deserialize(proof.serialize()).to_raw() == proof.to_raw()
    vk = params.get_vk()
    assert fb.VK.deserialize(vk.serialize()).to_raw() == vk.to_raw()
    inputs = fr_np([int(golden["inputs"][1], 16)])
    assert fb.verify(vk, proof, inputs) is True
    assert fb.verify(vk, proof, fr_np([int(golden["inputs"][1], 16) ^ 1])) is False
    tampered = bytearray(proof.to_raw())
    tampered[200] ^= 1
    assert fb.verify(vk, fb.Proof.from_raw(bytes(tampered)), inputs) is False
    with pytest.raises(fb.native.FbError) as e:   # reference: MalformedVerifyingKey -> panic
        fb.verify(vk, proof, fr_np([1, 2]))
    assert e.value.code == -7


def test_prove_finish_combines_partials(golden):
    """fb_prove_finish (host): partial sums from two base shards -> the golden proof."""
    import fawkes_crypto_b200 as fb
    from tests.dist_helpers import shard_partials
    pb = bytes.fromhex(golden["bellman_params_hex"])
    parts = b"".join(shard_partials(golden, rank, 2) for rank in range(2))
    r, s = fr_np([int(golden["r"], 16)])[0], fr_np([int(golden["s"], 16)])[0]
    out = np.zeros(256, dtype=np.uint8)
    pbuf = np.frombuffer(parts, dtype=np.uint8).copy()
    fb.native.check(fb.native.lib.fb_prove_finish(fb.native.ptr(pb), len(pb), pbuf.ctypes.data, 2, r.ctypes.data,
                                                  s.ctypes.data, out.ctypes.data))
    assert out.tobytes().hex() == golden["proof_raw_hex"]
