"""configs[1] on the GPU path: 256 proofs of the EdDSA-Poseidon signature circuit (c_eddsaposeidon_verify,
4,121 + 2 gates at the reference's commit, m = 2^13) built by oracle/frontend.py, one resident key, through
prove_batch (fb_prove_batch).  8 distinct signatures (keys, messages, witnesses) cycled over the 256 slots,
distinct r, s per proof.  Prints one JSON line."""
import json, os, random, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fawkes_crypto_b200 as fb
from oracle import bn254 as bn, codec, synth
from oracle import frontend as fe
from tests.util import fr_np

count, distinct = 256, 8
ctx = fb.Context(0)
rng = random.Random(2027)
jj, P = fe.JubJubBN256(), fe.PoseidonParams(4, 8, 54)
t = time.perf_counter()
cases = [fe.eddsa_circuit(rng.randrange(fe.FS), rng.randrange(bn.R), P, jj) for _ in range(distinct)]
t_front = (time.perf_counter() - t) / distinct
gates = cases[0][0]
raw = b"".join(codec.gate_borsh(g) for g in gates)
circ = fb.Circuit.from_raw_gates(raw, len(gates), 2, len(cases[0][2]))
td, r0, s0 = synth.synth_trapdoor(synth.SEED_BASE + 2)
t = time.perf_counter()
params = fb.setup(circ, ctx, trapdoor=[td.alpha, td.beta, td.gamma, td.delta, td.tau], gates_blob=codec.brotli_compress(raw))
t_setup = time.perf_counter() - t
t = time.perf_counter(); params.load(ctx); t_load = time.perf_counter() - t
wd = [(fr_np(inp), fr_np(aux)) for _, inp, aux in cases]
wit = [wd[i % distinct] for i in range(count)]
rs = [(r0 + 5 * i) % bn.R for i in range(count)]
ss = [(s0 + 9 * i) % bn.R for i in range(count)]
res = {}
ref = None
for slots in (1, 8):
    os.environ["FB_BATCH_SLOTS"] = str(slots)
    for rep in range(3):       # the first pass creates the slots (and captures the graph)
        t = time.perf_counter()
        out = fb.prove_batch(params, wit, rs, ss, ctx)
        dt = time.perf_counter() - t
    raws = [p.to_raw() for _, p in out]
    if ref is None:
        ref = raws
    assert raws == ref, f"batch with {slots} slots differs"
    res[f"slots_{slots}"] = {"batch_s": dt, "ms_per_proof": dt / count * 1e3, "proofs_per_s": count / dt}
t = time.perf_counter()
for i in range(20):
    _, single = fb.prove_with_rs(params, wit[i][0], wit[i][1], rs[i], ss[i], ctx)
    assert single.to_raw() == ref[i]
t_single = (time.perf_counter() - t) / 20
ok = all(fb.verify(params.get_vk(), out[i][1], out[i][0]) for i in (0, 1, 7, 100, 255))
bad = fb.verify(params.get_vk(), out[0][1], out[1][0])
print(json.dumps({"config": f"{count} proofs of the EdDSA-Poseidon circuit ({len(gates)} gates, {circ.shape()['nnz']} nnz, "
                            f"m = 2^{params.info()['log_m']}), {distinct} distinct signatures, one resident key",
                  "verified": bool(ok), "wrong_message_rejected": not bad, "single_prove_ms": t_single * 1e3,
                  "setup_s": t_setup, "key_load_s": t_load, "front_end_python_s_per_witness": t_front, **res}))
