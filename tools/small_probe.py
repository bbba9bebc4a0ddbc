"""Latency of small proves (cfg-1 / cfg-2 shapes: 2^13 and 2^12 rows) and a few sizes in between."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fawkes_crypto_b200 as fb
import bench
ctx = fb.Context(0)
for lg in [int(x) for x in os.environ.get("LOGS", "12,13,16,18").split(",")]:
    circ, params, tdi, setup_s = bench.make_case(fb, ctx, lg)
    wi, wa = circ.witness()
    for i in range(3):
        fb.groth16.prove_with_rs(params, wi, wa, tdi[5], tdi[6], ctx)
    t = time.perf_counter(); n = 20
    for i in range(n):
        inputs, proof = fb.groth16.prove_with_rs(params, wi, wa, tdi[5], tdi[6], ctx)
    dt = (time.perf_counter() - t) / n
    print(f"2^{lg}: {dt*1e3:.3f} ms/prove  stages {params.timings()}  verify {fb.verify(params.get_vk(), proof, inputs)}", flush=True)
    params.unload()
