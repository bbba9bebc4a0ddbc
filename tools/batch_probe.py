"""configs[1] shape: 256 proofs of a 2^12-domain circuit (3,862 rows, the EdDSA-Poseidon verify circuit's size;
synthetic rows -- its front end is not restated) on one resident key through fb_prove_batch, for several
numbers of in-flight slots.  Every proof is compared with the proof-by-proof result and a few are verified."""
import ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fawkes_crypto_b200 as fb

ctx = fb.Context(0)
lib = fb.native.lib
n_rows, count, seed = 3862, 256, 0xFA3CE50000 + 2
circ = fb.Circuit.synthetic(n_rows, seed)
td = np.zeros((7, 4), dtype=np.uint64)
fb.native.check(lib.fb_synth_trapdoor(seed, td.ctypes.data))
tdi = [fb.groth16.fr_unraw(x) for x in td]
params = fb.setup(circ, ctx, trapdoor=tdi[:5])
pk = params.load(ctx)
wi, wa = circ.witness()
sh = circ.shape()
rng = np.random.default_rng(3)
rs = rng.integers(0, 1 << 62, size=(2, count, 4), dtype=np.uint64); rs[:, :, 3] &= np.uint64((1 << 60) - 1)
ins = (C.c_void_p * count)(*[wi.ctypes.data] * count)
axs = (C.c_void_p * count)(*[wa.ctypes.data] * count)
res = {}
ref = None
for slots in (1, 2, 4, 8, 16):
    os.environ["FB_BATCH_SLOTS"] = str(slots)
    out = np.zeros((count, 256), dtype=np.uint8)
    for rep in range(2):   # first pass creates the slots
        t = time.perf_counter()
        fb.native.check(lib.fb_prove_batch(ctx.handle, pk, count, ins, sh["n_in"], axs, sh["n_aux"], rs[0].ctypes.data,
                                           rs[1].ctypes.data, out.ctypes.data))
        dt = time.perf_counter() - t
    if ref is None:
        ref = out.copy()
    assert np.array_equal(out, ref), f"batch with {slots} slots differs from the sequential proofs"
    res[f"slots_{slots}"] = {"batch_s": dt, "ms_per_proof": dt / count * 1e3, "proofs_per_s": count / dt}
    print(slots, res[f"slots_{slots}"], flush=True)
ok = all(fb.verify(params.get_vk(), fb.Proof.from_raw(ref[i].tobytes()), wi[1:]) for i in (0, 1, 100, 255))
print(json.dumps({"config": f"{count} proofs, synthetic {n_rows}-row circuit (m = 2^12), one resident key", "verified": bool(ok), **res}))
