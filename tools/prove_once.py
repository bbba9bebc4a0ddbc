"""Setup + a few proves of the synthetic 2^LOG circuit -- target for ncu launch lists / captures."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fawkes_crypto_b200 as fb
import bench
lg = int(os.environ.get("LOG", "20")); reps = int(os.environ.get("REPS", "2"))
ctx = fb.Context(0)
circ, params, tdi, setup_s = bench.make_case(fb, ctx, lg)
if os.environ.get("SERIAL", "0") == "1":
    fb.native.lib.fb_set_serial(1)
wi, wa = circ.witness()
for i in range(reps):
    t = time.time()
    inputs, proof = fb.groth16.prove_with_rs(params, wi, wa, tdi[5], tdi[6], ctx)
    print("prove", i, time.time() - t, params.timings(), flush=True)
print("verify", fb.verify(params.get_vk(), proof, inputs))
