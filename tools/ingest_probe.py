"""Gate-blob ingest: host parser (pk.cu) against the GPU ingest (ingest.cu) on random gate streams of
2^LOG gates (7 terms per gate, 37 B each -- the synthetic circuits' shape).  Prints one JSON line."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fawkes_crypto_b200 as fb
from tests.util import random_gate_blob

ctx = fb.Context(0)
res = {}
for lg in [int(x) for x in os.environ.get("LOGS", "18,20,22").split(",")]:
    n = 1 << lg
    raw = random_gate_blob(n, 2, n, seed=lg)
    r = {"blob_bytes": len(raw), "terms": 7 * n}
    for rep in range(2):
        t = time.perf_counter()
        dev = fb.Circuit.from_raw_gates(raw, n, 2, n, ctx=ctx)
        r["gpu_s"] = time.perf_counter() - t
        r["gpu_stage_ms"] = dev.ingest_ms
    if lg <= int(os.environ.get("HOST_MAX_LOG", "20")):
        t = time.perf_counter()
        host = fb.Circuit.from_raw_gates(raw, n, 2, n)
        r["host_s"] = time.perf_counter() - t
        r["speedup"] = r["host_s"] / r["gpu_s"]
        assert host.shape() == dev.shape()
    r["terms_per_s_gpu"] = 7 * n / r["gpu_s"]
    res[f"2^{lg}"] = r
    print(lg, r, file=sys.stderr, flush=True)
print(json.dumps({"config": "random borsh gate streams, 3+3+1 terms per gate; GPU = framing walk on the host + upload + "
                            "kernels + copy back of the CSR", **res}))
