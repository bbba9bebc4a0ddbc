"""End-to-end parity: setup / prove / verify through the C ABI vs the oracle (byte-exact)."""
import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import codec
from oracle import groth16 as og
from oracle import synth
from tests.util import fr_np, fr_list

pytestmark = pytest.mark.gpu


def oracle_case(n_rows, seed):
    gates, inp, aux = synth.synth_circuit(n_rows, seed)
    td, r, s = synth.synth_trapdoor(seed)
    P = og.setup(gates, 2, len(aux), td)
    return gates, inp, aux, td, r, s, P


@pytest.mark.parametrize("tables", [0, 1], ids=["plain_msm", "window_tables"])
@pytest.mark.parametrize("n_rows", [3, 17, 64, 700])
def test_prove_matches_oracle_bytes(ctx, n_rows, tables):
    import fawkes_crypto_b200 as fb
    fb.native.lib.fb_set_msm_tables(tables)   # keys loaded below use / do not use the MSM window tables
    seed = synth.SEED_BASE + 100 + n_rows
    gates, inp, aux, td, r, s, P = oracle_case(n_rows, seed)
    ref_proof, ref_h = og.prove(P, gates, inp, aux, r, s, return_h=True)
    raw = b"".join(codec.gate_borsh(g) for g in gates)
    params = fb.Parameters(codec.bellman_params_bytes(P), len(gates), codec.brotli_compress(raw))
    # through the reference's own byte framing (Parameters::write -> read)
    params = fb.Parameters.read(params.write())
    inputs, proof, h = fb.groth16.prove_with_rs(params, fr_np(inp), fr_np(aux), r, s, ctx, return_h=True)
    assert fr_list(h) == ref_h
    assert proof.to_raw() == codec.proof_raw(ref_proof)
    assert proof.serialize() == codec.proof_borsh(ref_proof)
    assert fr_list(inputs) == inp[1:]
    assert fb.verify(params.get_vk(), proof, inputs)
    assert og.verify(P.vk, codec.proof_unraw(proof.to_raw()), inp[1:])
    # tampered public input must fail
    bad = fr_np([(inp[1] + 1) % bn.R])
    assert not fb.verify(params.get_vk(), proof, bad)
    params.unload()
    fb.native.lib.fb_set_msm_tables(-1)


@pytest.mark.parametrize("n_rows", [5, 64, 300])
def test_setup_matches_oracle_bytes(ctx, n_rows):
    import fawkes_crypto_b200 as fb
    seed = synth.SEED_BASE + 200 + n_rows
    gates, inp, aux, td, r, s, P = oracle_case(n_rows, seed)
    circ = fb.Circuit.synthetic(n_rows, seed)
    params = fb.setup(circ, ctx, trapdoor=[td.alpha, td.beta, td.gamma, td.delta, td.tau])
    assert bytes(params.bellman_bytes) == codec.bellman_params_bytes(P)


def test_prove_synthetic_2_16_scalar_identity_and_verify(ctx):
    """2^16 rows: too big for the Python prover, so check (i) the trapdoor scalar identity
    A == (alpha + A(tau) + r delta) G1 etc. with the oracle's O(n) scalar side, (ii) verify."""
    import fawkes_crypto_b200 as fb
    n_rows = 1 << 16
    seed = synth.SEED_BASE + 16
    circ = fb.Circuit.synthetic(n_rows, seed)
    tdn = np.zeros((7, 4), dtype=np.uint64)
    fb.native.check(fb.native.lib.fb_synth_trapdoor(seed, tdn.ctypes.data))
    td = fr_list(tdn)
    params = fb.setup(circ, ctx, trapdoor=td[:5])
    wi, wa = circ.witness()
    inputs, proof = fb.groth16.prove_with_rs(params, wi, wa, td[5], td[6], ctx)
    assert fb.verify(params.get_vk(), proof, inputs)
    gates, inp, aux = synth.synth_circuit(n_rows, seed)
    assert fr_list(wa) == aux
    P = og.setup(gates, 2, len(aux), og.Trapdoor(*td[:5]), want_points=False)
    A, B, C = og.prove_scalar_side(P, gates, inp, aux, td[5], td[6])
    pr = codec.proof_unraw(proof.to_raw())
    assert pr.a == bn.pt_mul(bn.OPS1, bn.G1_GEN, A)
    assert pr.b == bn.pt_mul(bn.OPS2, bn.G2_GEN, B)
    assert pr.c == bn.pt_mul(bn.OPS1, bn.G1_GEN, C)
    params.unload()


def test_error_paths(ctx):
    import fawkes_crypto_b200 as fb
    seed = synth.SEED_BASE + 300
    gates, inp, aux, td, r, s, P = oracle_case(20, seed)
    raw = b"".join(codec.gate_borsh(g) for g in gates)
    good = codec.bellman_params_bytes(P)
    # wrong witness length
    params = fb.Parameters(good, len(gates), codec.brotli_compress(raw))
    with pytest.raises(fb.native.FbError) as e:
        fb.groth16.prove_with_rs(params, fr_np(inp), fr_np(aux[:-1]), r, s, ctx)
    assert e.value.code == -1
    params.unload()
    # truncated parameters
    with pytest.raises(fb.native.FbError) as e:
        fb.Parameters(good[:-10], len(gates), codec.brotli_compress(raw)).load(ctx)
    assert e.value.code == -3
    # point not on curve (checked = True)
    bad = bytearray(good)
    bad[64 + 63] ^= 1   # beta_g1.y
    with pytest.raises(fb.native.FbError):
        fb.Parameters(bytes(bad), len(gates), codec.brotli_compress(raw)).load(ctx)
    # gate count mismatch
    with pytest.raises(fb.native.FbError) as e:
        fb.Parameters(good, len(gates) + 1, codec.brotli_compress(raw)).load(ctx)
    assert e.value.code == -3


def test_parameters_read_flags(ctx):
    """Parameters::read(reader, disallow_points_at_infinity, checked) (mod.rs:159-175): `checked` adds the curve and
    the r-torsion test to every query point (the subgroup test matters for G2 only), `disallow_points_at_infinity`
    rejects the infinity encoding, and the decoder always rejects a dirty infinity encoding."""
    import struct
    import fawkes_crypto_b200 as fb
    from tests.test_abi import g2_point_outside_the_subgroup
    seed = synth.SEED_BASE + 301
    gates, inp, aux, td, r, s, P = oracle_case(20, seed)
    raw = b"".join(codec.gate_borsh(g) for g in gates)
    blob = codec.brotli_compress(raw)
    good = codec.bellman_params_bytes(P)
    framed = lambda pb: fb.Parameters(pb, len(gates), blob).write()
    # section offsets of the bellman body
    n_ic = struct.unpack_from(">I", good, 576)[0]
    pos = 580 + 64 * n_ic
    off = {}
    for name, sz in (("h", 64), ("l", 64), ("a", 64), ("b_g1", 64), ("b_g2", 128)):
        n = struct.unpack_from(">I", good, pos)[0]
        off[name] = pos + 4
        pos += 4 + n * sz
    # 1. a b_g2 query point on the twist curve but outside the subgroup: only a checked read notices
    bad = bytearray(good)
    bad[off["b_g2"]:off["b_g2"] + 128] = codec.g2_uncompressed(g2_point_outside_the_subgroup(3))
    with pytest.raises(fb.native.FbError) as e:
        fb.Parameters.read(framed(bytes(bad)), False, True).load(ctx)
    assert e.value.code == -3 and "r-torsion" in str(e.value)
    p_unchecked = fb.Parameters.read(framed(bytes(bad)), False, False)
    p_unchecked.load(ctx)
    p_unchecked.unload()
    # 2. a point at infinity in the h query: fine by default, an error with disallow_points_at_infinity
    inf = bytearray(good)
    inf[off["h"]:off["h"] + 64] = bytes([0x40]) + bytes(63)
    p_inf = fb.Parameters.read(framed(bytes(inf)), False, True)
    p_inf.load(ctx)
    p_inf.unload()
    with pytest.raises(fb.native.FbError) as e:
        fb.Parameters.read(framed(bytes(inf)), True, True).load(ctx)
    assert e.value.code == -3 and "infinity" in str(e.value)
    # 3. infinity flag with other bits set: never accepted
    inf[off["h"] + 40] = 7
    with pytest.raises(fb.native.FbError) as e:
        fb.Parameters.read(framed(bytes(inf)), False, False).load(ctx)
    assert e.value.code == -3
    # 4. an l query point off the curve: only a checked read notices
    offc = bytearray(good)
    offc[off["l"] + 63] ^= 1
    with pytest.raises(fb.native.FbError):
        fb.Parameters.read(framed(bytes(offc)), False, True).load(ctx)
    # 5. one resident key per Parameters: another context / shard needs unload() first
    params = fb.Parameters.read(framed(good))
    params.load(ctx)
    with pytest.raises(ValueError):
        params.load(ctx, shard=1, nshards=2)
    inputs, proof = fb.groth16.prove_with_rs(params, fr_np(inp), fr_np(aux), r, s, ctx)
    assert proof.to_raw() == codec.proof_raw(og.prove(P, gates, inp, aux, r, s))
    params.unload()


def test_prove_rejects_a_key_of_another_context(ctx):
    import ctypes as C
    import fawkes_crypto_b200 as fb
    seed = synth.SEED_BASE + 302
    gates, inp, aux, td, r, s, P = oracle_case(12, seed)
    raw = b"".join(codec.gate_borsh(g) for g in gates)
    params = fb.Parameters(codec.bellman_params_bytes(P), len(gates), codec.brotli_compress(raw))
    pk = params.load(ctx)
    other = fb.Context(ctx.device)
    vi, va = fr_np(inp), fr_np(aux)
    ra, sa = fr_np([r])[0], fr_np([s])[0]
    out = np.zeros(256, dtype=np.uint8)
    rc = fb.native.lib.fb_prove(other.handle, pk, vi.ctypes.data, 2, va.ctypes.data, va.shape[0], ra.ctypes.data,
                                sa.ctypes.data, out.ctypes.data, None)
    assert rc == -1 and "different fb_ctx" in fb.native.last_error()
    other.close()
    params.unload()


def test_corrupt_gate_stream_is_an_error(ctx):
    """A gate blob that is not valid brotli, is cut short, or expands beyond what the announced gate count can
    hold is FB_ERR_FORMAT at key load (never a silently shorter circuit, never unbounded growth)."""
    import fawkes_crypto_b200 as fb
    gates, inp, aux = synth.synth_circuit(200, 11)
    raw = b"".join(codec.gate_borsh(g) for g in gates)
    blob = codec.brotli_compress(raw)
    for bad in (blob[:len(blob) // 2], b"\xff" * 64 + blob[64:]):
        with pytest.raises(fb.native.FbError) as e:
            fb.Circuit.from_gates_blob(bad, len(gates), 2, len(aux), ctx)
        assert e.value.code == -3
        with pytest.raises(fb.native.FbError) as e:
            fb.Circuit.from_gates_blob(bad, len(gates), 2, len(aux))
        assert e.value.code == -3
    bomb = codec.brotli_compress(bytes(64 << 20))       # 64 MiB of zeros in a few hundred bytes
    with pytest.raises(fb.native.FbError) as e:
        fb.Circuit.from_gates_blob(bomb, 3, 2, 5)
    assert e.value.code == -3 and "expands beyond" in str(e.value)


def test_pageable_witness_takes_the_staged_upload(ctx):
    """A witness in plain (pageable) host memory -- what a Rust Vec<Num<Fr>> is -- above 4 MiB goes through the
    library's pinned staging ring (api.cu: upload_host, four copy threads); the proof must equal the one made from
    pinned memory and from a device-resident witness, and repeated calls must reuse the ring correctly."""
    import torch
    import fawkes_crypto_b200 as fb
    import bench
    lib = fb.native.lib
    circ, params, tdi, _ = bench.make_case(fb, ctx, 18)          # 262k aux values = 8.4 MB
    pk = params.load(ctx, checked=False)
    wi, wa = circ.witness()
    n_in, n_aux = wi.shape[0], wa.shape[0]
    r, s = fb.groth16.fr_raw(tdi[5]), fb.groth16.fr_raw(tdi[6])
    page = np.empty((n_in + n_aux, 4), dtype=np.uint64)          # pageable
    page[:n_in], page[n_in:] = wi, wa
    pinned = torch.empty((n_in + n_aux, 4), dtype=torch.int64).pin_memory()
    pinned.numpy().view(np.uint64)[:] = page
    outs = []
    for src in (page, pinned.numpy().view(np.uint64), page, page):
        out = np.zeros(256, dtype=np.uint8)
        fb.native.check(lib.fb_prove(ctx.handle, pk, src.ctypes.data, n_in, src[n_in:].ctypes.data, n_aux, r.ctypes.data,
                                     s.ctypes.data, out.ctypes.data, None))
        outs.append(out.tobytes())
    dev = pinned.cuda()
    out = np.zeros(256, dtype=np.uint8)
    fb.native.check(lib.fb_prove_device(ctx.handle, pk, dev.data_ptr(), r.ctypes.data, s.ctypes.data, out.ctypes.data))
    assert all(o == out.tobytes() for o in outs)
    assert fb.verify(params.get_vk(), fb.Proof.from_raw(outs[0]), wi[1:])
    # a different witness in the same pageable buffer must give a different (and still valid) statement: the ring is
    # not serving stale chunks
    page[n_in + 5, 0] ^= np.uint64(1)
    out2 = np.zeros(256, dtype=np.uint8)
    fb.native.check(lib.fb_prove(ctx.handle, pk, page.ctypes.data, n_in, page[n_in:].ctypes.data, n_aux, r.ctypes.data,
                                 s.ctypes.data, out2.ctypes.data, None))
    assert out2.tobytes() != outs[0]
    params.unload()


def test_stream_close_with_queued_proofs_does_not_hang(ctx):
    """fb_stream_close while proofs are still queued: the ones not started are dropped with an error result, the call
    returns, and the key is usable afterwards."""
    import fawkes_crypto_b200 as fb
    seed = synth.SEED_BASE + 6100
    gates, inp, aux = synth.synth_circuit(300, seed)
    td, r0, s0 = synth.synth_trapdoor(seed)
    P = og.setup(gates, 2, len(aux), td)
    raw = b"".join(codec.gate_borsh(g) for g in gates)
    params = fb.Parameters(codec.bellman_params_bytes(P), len(gates), codec.brotli_compress(raw))
    wi, wa = fr_np(inp), fr_np(aux)
    st = fb.ProveStream(params, ctx, depth=16)
    for i in range(16):
        st.submit(wi, wa, (r0 + i) % bn.R, (s0 + i) % bn.R)
    st.close()
    _, proof = fb.groth16.prove_with_rs(params, wi, wa, r0, s0, ctx)
    assert proof.to_raw() == codec.proof_raw(og.prove(P, gates, inp, aux, r0, s0))
    params.unload()


def test_golden_fixture_through_blob_path(ctx):
    """Committed golden vector (tests/golden, made by tools/gen_golden.py from the oracle): raw
    Parameters bytes + brotli gate blob -> fb_pk_load -> proof bytes."""
    import json
    import os
    import ctypes as C
    import fawkes_crypto_b200 as fb
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "synth_rows40.json")))
    pb = bytes.fromhex(g["bellman_params_hex"])
    blob = codec.brotli_compress(bytes.fromhex(g["gates_raw_hex"]))
    pk = C.c_void_p()
    fb.native.check(fb.native.lib.fb_pk_load(ctx.handle, fb.native.ptr(pb), len(pb), fb.native.ptr(blob), len(blob),
                                             38, 1, C.byref(pk)))
    inp = fr_np([int(x, 16) for x in g["inputs"]])
    aux = fr_np([int(x, 16) for x in g["aux"]])
    r, s = fr_np([int(g["r"], 16)])[0], fr_np([int(g["s"], 16)])[0]
    out = np.zeros(256, dtype=np.uint8)
    fb.native.check(fb.native.lib.fb_prove(ctx.handle, pk, inp.ctypes.data, 2, aux.ctypes.data, aux.shape[0],
                                           r.ctypes.data, s.ctypes.data, out.ctypes.data, None))
    assert out.tobytes().hex() == g["proof_raw_hex"]
    # the device-resident entry point gives the same bytes
    import torch
    w = torch.from_numpy(np.concatenate([inp, aux]).view(np.int64)).cuda()
    out2 = np.zeros(256, dtype=np.uint8)
    fb.native.check(fb.native.lib.fb_prove_device(ctx.handle, pk, w.data_ptr(), r.ctypes.data, s.ctypes.data,
                                                  out2.ctypes.data))
    assert out2.tobytes() == out.tobytes()
    fb.native.lib.fb_pk_free(pk)


@pytest.mark.parametrize("nshards", [2, 3])
def test_sharded_prove_equals_unsharded(ctx, nshards):
    """Base-index shards (the multi-GPU path) proved one after the other on one GPU and combined
    with fb_prove_finish give the same proof bytes as the unsharded prover."""
    import ctypes as C
    import fawkes_crypto_b200 as fb
    seed = synth.SEED_BASE + 4000 + nshards
    n_rows = 3000
    circ = fb.Circuit.synthetic(n_rows, seed)
    tdn = np.zeros((7, 4), dtype=np.uint64)
    fb.native.check(fb.native.lib.fb_synth_trapdoor(seed, tdn.ctypes.data))
    td = fr_list(tdn)
    params = fb.setup(circ, ctx, trapdoor=td[:5])
    wi, wa = circ.witness()
    _, ref = fb.groth16.prove_with_rs(params, wi, wa, td[5], td[6], ctx)
    params.unload()
    parts = np.zeros((nshards, 640), dtype=np.uint8)
    for sh in range(nshards):
        # every other shard loads its key from the sharded setup (fb_setup_shard: only this shard's query points are
        # generated, everything else is the point at infinity), the rest from the full Parameters
        pb = params.bellman_bytes if sh % 2 else fb.setup(circ, ctx, trapdoor=td[:5], shard=sh, nshards=nshards).bellman_bytes
        assert len(pb) == len(params.bellman_bytes)
        pk = C.c_void_p()
        fb.native.check(fb.native.lib.fb_pk_load_shard(ctx.handle, fb.native.ptr(pb), len(pb), circ.handle, 1, sh,
                                                       nshards, C.byref(pk)))
        fb.native.check(fb.native.lib.fb_prove_partial(ctx.handle, pk, wi.ctypes.data, wi.shape[0], wa.ctypes.data,
                                                       wa.shape[0], parts[sh].ctypes.data))
        fb.native.lib.fb_pk_free(pk)
    r, s = fr_np([td[5]])[0], fr_np([td[6]])[0]
    out = np.zeros(256, dtype=np.uint8)
    pb = params.bellman_bytes
    fb.native.check(fb.native.lib.fb_prove_finish(fb.native.ptr(pb), len(pb), parts.ctypes.data, nshards,
                                                  r.ctypes.data, s.ctypes.data, out.ctypes.data))
    assert out.tobytes() == ref.to_raw()
    assert fb.verify(params.get_vk(), fb.Proof.from_raw(out.tobytes()), wi[1:])


def test_witness_with_zeros_and_ones(ctx):
    """Real witnesses are full of 0/1: a circuit of boolean constraints b*(b-1)=0 plus products,
    through the gate-blob path (exercises skewed buckets and cidx = -1)."""
    import fawkes_crypto_b200 as fb
    import random
    rng = random.Random(5)
    n_bits = 600
    aux = [rng.choice([0, 1]) for _ in range(n_bits)]
    gates = []
    AUX, INPUT = og.AUX, og.INPUT
    for i in range(n_bits):      # b * (b - 1) = 0
        gates.append(([(1, (AUX, i))], [(1, (AUX, i)), (bn.R - 1, (INPUT, 0))], []))
    acc = sum(aux) % bn.R        # public: number of ones, as sum * 1 = input_1
    gates.append(([(1, (AUX, i)) for i in range(n_bits)], [(1, (INPUT, 0))], [(1, (INPUT, 1))]))
    inp = [1, acc]
    td, r, s = synth.synth_trapdoor(12345)
    P = og.setup(gates, 2, len(aux), td)
    ref = og.prove(P, gates, inp, aux, r, s)
    raw = b"".join(codec.gate_borsh(g) for g in gates)
    params = fb.Parameters(codec.bellman_params_bytes(P), len(gates), codec.brotli_compress(raw))
    inputs, proof = fb.groth16.prove_with_rs(params, fr_np(inp), fr_np(aux), r, s, ctx)
    assert proof.to_raw() == codec.proof_raw(ref)
    assert fb.verify(params.get_vk(), proof, inputs)
    params.unload()


@pytest.mark.parametrize("mode,chunk,tables", [("batched", 64, 1), ("batched", 3, 1), ("batched", 2, 0), ("slots", 0, 1)])
def test_prove_batch_matches_single_proofs(ctx, mode, chunk, tables):
    """fb_prove_batch (configs[1] shape: many proofs on one resident key) == the oracle's proof of each witness.
    batched = one set of launches per chunk of proofs, buckets keyed by (proof, digit) -- with chunks that divide
    the batch, chunks with a remainder, and without window tables (bucket sets per (proof, window)); slots = the
    older scheme (independent proves in flight).  DIFFERENT witnesses per proof: a second circuit input changes
    every aux value downstream of it."""
    import ctypes as C
    import os
    import fawkes_crypto_b200 as fb
    seed = synth.SEED_BASE + 5000
    n_rows, count = 500, 7
    gates, inp, aux = synth.synth_circuit(n_rows, seed)
    td, r0, s0 = synth.synth_trapdoor(seed)
    P = og.setup(gates, 2, len(aux), td)
    raw = b"".join(codec.gate_borsh(g) for g in gates)
    # witnesses: re-evaluate the circuit from different initial aux values (rows 1.. define aux[16..] as products)
    wits = []
    for i in range(count):
        a = list(aux[:synth.N_INIT_AUX])
        a[3] = (a[3] + 1000 * i) % bn.R
        ins_i = [1, a[0]]
        for g in gates[1:]:
            val = lambda t: ins_i[t[1]] if t[0] == og.INPUT else a[t[1]]
            ea = sum(c * val(t) for c, t in g[0]) % bn.R
            eb = sum(c * val(t) for c, t in g[1]) % bn.R
            a.append(ea * eb % bn.R)
        wits.append((ins_i, a))
    assert wits[0][1] == aux and wits[1][1] != aux
    fb.native.lib.fb_set_msm_tables(tables)
    os.environ["FB_BATCH_MODE"] = mode
    if chunk:
        os.environ["FB_BATCH_P"] = str(chunk)
    try:
        params = fb.Parameters(codec.bellman_params_bytes(P), len(gates), codec.brotli_compress(raw))
        pk = params.load(ctx)
        rs = [((r0 + 7 * i) % bn.R, (s0 + 11 * i) % bn.R) for i in range(count)]
        wis = [fr_np(w[0]) for w in wits]
        was = [fr_np(w[1]) for w in wits]
        ins = (C.c_void_p * count)(*[x.ctypes.data for x in wis])
        axs = (C.c_void_p * count)(*[x.ctypes.data for x in was])
        ra, sa = fr_np([x for x, _ in rs]), fr_np([y for _, y in rs])
        out = np.zeros((count, 256), dtype=np.uint8)
        for rep in range(2):      # the second call reuses the batch workspaces
            out[:] = 0
            fb.native.check(fb.native.lib.fb_prove_batch(ctx.handle, pk, count, ins, 2, axs, len(aux), ra.ctypes.data,
                                                         sa.ctypes.data, out.ctypes.data))
            for i, (r, s) in enumerate(rs):
                ref = og.prove(P, gates, wits[i][0], wits[i][1], r, s)
                assert out[i].tobytes() == codec.proof_raw(ref), (rep, i)
        # a wrong-shaped witness is refused
        rc = fb.native.lib.fb_prove_batch(ctx.handle, pk, count, ins, 2, axs, len(aux) - 1, ra.ctypes.data, sa.ctypes.data,
                                          out.ctypes.data)
        assert rc == -1
        params.unload()
    finally:
        fb.native.lib.fb_set_msm_tables(-1)
        os.environ.pop("FB_BATCH_MODE", None)
        os.environ.pop("FB_BATCH_P", None)


def test_prove_stream_matches_single_proofs(ctx):
    """fb_stream_* (witness k+1 is prepared while proof k runs): every proof equals the oracle's, tickets can be
    collected out of order, a ticket is good once, a wrong-shaped witness is refused at submit."""
    import fawkes_crypto_b200 as fb
    seed = synth.SEED_BASE + 6000
    gates, inp, aux = synth.synth_circuit(400, seed)
    td, r0, s0 = synth.synth_trapdoor(seed)
    P = og.setup(gates, 2, len(aux), td)
    raw = b"".join(codec.gate_borsh(g) for g in gates)
    params = fb.Parameters(codec.bellman_params_bytes(P), len(gates), codec.brotli_compress(raw))
    wi, wa = fr_np(inp), fr_np(aux)
    count = 12
    rs = [((r0 + 3 * i) % bn.R, (s0 + 13 * i) % bn.R) for i in range(count)]
    with fb.ProveStream(params, ctx, depth=3) as st:       # depth < count: submit has to wait for space
        tickets = []
        for r, s in rs:
            scratch_i, scratch_a = wi.copy(), wa.copy()
            tickets.append(st.submit(scratch_i, scratch_a, r, s))
            scratch_i[:] = 0                                 # the caller's buffers are free again after submit
            scratch_a[:] = 0
        assert len(set(tickets)) == count
        got = {}
        for i in reversed(range(count)):
            inputs, proof = st.wait(tickets[i])
            assert fr_list(inputs) == inp[1:]
            got[i] = proof.to_raw()
        with pytest.raises(fb.native.FbError):
            st.wait(tickets[0])
        with pytest.raises(fb.native.FbError):
            st.submit(wi, wa[:-1], r0, s0)
    for i in (0, 5, count - 1):
        assert got[i] == codec.proof_raw(og.prove(P, gates, inp, aux, rs[i][0], rs[i][1])), i
    # the key is usable again after the stream is closed
    _, proof = fb.groth16.prove_with_rs(params, wi, wa, rs[1][0], rs[1][1], ctx)
    assert proof.to_raw() == got[1]
    params.unload()


def test_prove_2_20_bytes_equal_cpp_oracle(ctx):
    """Full BASELINE size (configs[2], 2^20 rows): the GPU proof is byte-identical to the C++ CPU
    restatement (itself byte-checked against the Python oracle in tests/test_oracle.py) and verifies."""
    import fawkes_crypto_b200 as fb
    import bench
    from oracle import cpu
    import hashlib
    circ, params, tdi, _ = bench.make_case(fb, ctx, 20)
    wi, wa = circ.witness()
    inputs, proof = fb.groth16.prove_with_rs(params, wi, wa, tdi[5], tdi[6], ctx)
    assert fb.verify(params.get_vk(), proof, inputs)
    # the oracle's own circuit and witness (oracle/cpu_setup.cpp), the keys fb_setup produced
    seed = bench.SEED_BASE + bench.cfg_number(20)
    ccirc = cpu.Circuit.synthetic(1 << 20, seed)
    assert np.array_equal(ccirc.aux, wa) and np.array_equal(ccirc.inputs, wi)
    ctd = cpu.synth_trapdoor(seed)
    cproof, _, _ = cpu.prove_circuit(params.bellman_bytes, ccirc, ctd[5], ctd[6], cpu.hw_threads())
    assert proof.to_raw() == cproof
    # committed golden (tools/gen_golden_synth.py: CPU circuit + CPU setup + CPU prove)
    g = bench.golden_synth(20)
    assert hashlib.sha256(proof.to_raw()).hexdigest() == g["proof_sha256"]
    assert hashlib.sha256(params.bellman_bytes).hexdigest() == g["params_sha256"]    # fb_setup == CPU setup, 335 MB
    params.unload()


@pytest.mark.parametrize("log_rows", [12, 16])
def test_gpu_setup_and_prove_equal_cpu_golden_small(ctx, log_rows):
    """tests/golden/synth_proofs.json at 2^12 and 2^16 rows: GPU setup digest and GPU proof bytes equal the CPU chain's."""
    import hashlib
    import fawkes_crypto_b200 as fb
    import bench
    circ, params, tdi, _ = bench.make_case(fb, ctx, log_rows)
    wi, wa = circ.witness()
    _, proof = fb.groth16.prove_with_rs(params, wi, wa, tdi[5], tdi[6], ctx)
    g = bench.golden_synth(log_rows)
    assert proof.to_raw().hex() == g["proof_raw_hex"]
    assert hashlib.sha256(params.bellman_bytes).hexdigest() == g["params_sha256"]
    params.unload()


def test_cfg1_poseidon_merkle_setup_prove_verify(ctx):
    """configs[0]: the reference's own hot-path test (tests/bellman_groth16.rs:19-47) -- Poseidon Merkle
    proof of depth 32, 7,328 + 34 gates, rows with up to ~55 terms -- built by the front-end restatement
    (oracle/frontend.py), then setup + prove + verify through the C ABI with a fixed trapdoor and r, s.
    The proof must equal the C++ CPU restatement's byte for byte, go through the reference's byte framing
    (Parameters::write/read, brotli gate blob) unchanged, and verify; a wrong root must not."""
    import random
    import fawkes_crypto_b200 as fb
    import bench
    from oracle import cpu
    from oracle import frontend as fe
    rng = random.Random(2026)
    leaf = rng.randrange(bn.R)
    sibling = [rng.randrange(bn.R) for _ in range(32)]
    path = [rng.random() < 0.5 for _ in range(32)]
    gates, inp, aux = fe.merkle_circuit(leaf, sibling, path)
    assert len(gates) == 7362 and len(aux) == 7394
    raw = b"".join(codec.gate_borsh(g) for g in gates)
    circ = fb.Circuit.from_raw_gates(raw, len(gates), 2, len(aux))
    assert circ.shape()["n_gates"] == 7362
    td, r, s = synth.synth_trapdoor(synth.SEED_BASE + 1)
    tdi = [td.alpha, td.beta, td.gamma, td.delta, td.tau]
    params = fb.setup(circ, ctx, trapdoor=tdi, gates_blob=codec.brotli_compress(raw))
    params = fb.Parameters.read(params.write())            # reference framing, gate blob included
    wi, wa = fr_np(inp), fr_np(aux)
    inputs, proof = fb.groth16.prove_with_rs(params, wi, wa, r, s, ctx)
    assert params.info()["log_m"] == 13 and params.info()["n_gates"] == 7362
    assert fr_list(inputs) == inp[1:]
    assert fb.verify(params.get_vk(), proof, inputs)
    assert not fb.verify(params.get_vk(), proof, fr_np([(inp[1] + 1) % bn.R]))
    # byte parity with the CPU restatement on the same key and witness
    sh = circ.shape()
    rp, cl, cf = bench.expand_csr(fb, circ)
    ref, _, _ = cpu.prove(params.bellman_bytes, sh["n_gates"], sh["n_in"], sh["n_aux"], rp, cl, cf, wi, wa,
                          fb.groth16.fr_raw(r), fb.groth16.fr_raw(s), 4)
    assert proof.to_raw() == ref
    params.unload()


def test_cfg2_eddsa_poseidon_batch_setup_prove_verify(ctx):
    """configs[1]: the EdDSA-Poseidon signature circuit (circuit/eddsaposeidon.rs:16-47; 4,121 + 2 gates at
    the reference's commit, m = 2^13) built by the front-end restatement, one resident key, several
    signatures (different keys and messages -> different witnesses and public inputs) proved as one batch.
    Every batch proof equals the proof-by-proof result and the C++ CPU restatement byte for byte and
    verifies against its own message only."""
    import random
    import fawkes_crypto_b200 as fb
    import bench
    from oracle import cpu
    from oracle import frontend as fe
    rng = random.Random(2027)
    jj, P = fe.JubJubBN256(), fe.PoseidonParams(4, 8, 54)
    count = 6
    cases = [fe.eddsa_circuit(rng.randrange(fe.FS), rng.randrange(bn.R), P, jj) for _ in range(count)]
    gates = cases[0][0]
    assert len(gates) == 4123 and all(c[0] == gates for c in cases)
    n_aux = len(cases[0][2])
    raw = b"".join(codec.gate_borsh(g) for g in gates)
    circ = fb.Circuit.from_raw_gates(raw, len(gates), 2, n_aux)
    td, r0, s0 = synth.synth_trapdoor(synth.SEED_BASE + 2)
    params = fb.setup(circ, ctx, trapdoor=[td.alpha, td.beta, td.gamma, td.delta, td.tau],
                      gates_blob=codec.brotli_compress(raw))
    params = fb.Parameters.read(params.write())
    wit = [(fr_np(inp), fr_np(aux)) for _, inp, aux in cases]
    rs = [(r0 + 5 * i) % bn.R for i in range(count)]
    ss = [(s0 + 9 * i) % bn.R for i in range(count)]
    batch = fb.prove_batch(params, wit, rs, ss, ctx)
    assert params.info()["log_m"] == 13
    sh = circ.shape()
    rp, cl, cf = bench.expand_csr(fb, circ)
    for i, (inputs, proof) in enumerate(batch):
        assert fr_list(inputs) == cases[i][1][1:]
        _, single = fb.prove_with_rs(params, wit[i][0], wit[i][1], rs[i], ss[i], ctx)
        assert single.to_raw() == proof.to_raw(), i
        assert fb.verify(params.get_vk(), proof, inputs)
        assert not fb.verify(params.get_vk(), proof, batch[(i + 1) % count][0])
        if i < 2:
            ref, _, _ = cpu.prove(params.bellman_bytes, sh["n_gates"], sh["n_in"], sh["n_aux"], rp, cl, cf,
                                  wit[i][0], wit[i][1], fb.groth16.fr_raw(rs[i]), fb.groth16.fr_raw(ss[i]), 4)
            assert proof.to_raw() == ref, i
    params.unload()


def test_prove_2_24_full_size_verifies(ctx):
    """BASELINE configs[3] at full size (2^24 rows, 75 GB resident key with window tables): the proof is BYTE-IDENTICAL to the CPU oracle's
    (committed golden), passes the pairing check (a size-independent end-to-end property: A, B and C are right only if all five
    MSMs over 16.7 M points and the seven 2^24-point transforms are), the public input is echoed, a
    tampered input fails, and proving twice gives the same bytes."""
    import fawkes_crypto_b200 as fb
    import bench
    circ, params, tdi, _ = bench.make_case(fb, ctx, 24)
    wi, wa = circ.witness()
    inputs, proof = fb.groth16.prove_with_rs(params, wi, wa, tdi[5], tdi[6], ctx)
    assert params.info()["log_m"] == 24 and params.info()["msm_tables"] == 1
    assert np.array_equal(np.asarray(inputs), wi[1:])
    assert fb.verify(params.get_vk(), proof, inputs)
    bad = np.array(wi[1:], copy=True)
    bad[0, 0] ^= np.uint64(1)
    assert not fb.verify(params.get_vk(), proof, bad)
    _, proof2 = fb.groth16.prove_with_rs(params, wi, wa, tdi[5], tdi[6], ctx)
    assert proof2.to_raw() == proof.to_raw()
    # byte parity at the full size: the committed golden is the CPU oracle's proof of this circuit, witness, r, s
    # under the CPU oracle's own setup (tools/gen_golden_synth.py, ~10 CPU-minutes); the digest of the 5.4 GB
    # Parameters byte string says fb_setup reproduced every one of the 84 M key points
    import hashlib
    g = bench.golden_synth(24)
    assert proof.to_raw().hex() == g["proof_raw_hex"]
    assert hashlib.sha256(params.bellman_bytes).hexdigest() == g["params_sha256"]
    params.unload()
