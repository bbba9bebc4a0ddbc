"""One-shot GPU probe: IMAD / Fr-mul peaks, H-pipeline and MSM timings.  Writes gpurun_out/probe.json."""
import json
import os
import sys
import time
import ctypes as C

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fawkes_crypto_b200 as fb  # noqa: E402

lib = fb.native.lib
ctx = fb.Context(0)
out = {}
d = C.c_double()
fb.native.check(lib.fb_probe_imad(ctx.handle, C.byref(d)))
out["imad_wide_mac_per_s"] = d.value
fb.native.check(lib.fb_probe_fr_mul(ctx.handle, C.byref(d)))
out["fr_mul_per_s"] = d.value
print(out, flush=True)

rng = np.random.default_rng(1)


def rand_fr(n):
    x = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    x[:, 3] &= np.uint64((1 << 60) - 1)
    return x


logs = [int(x) for x in os.environ.get("PROBE_LOGS", "16,20").split(",")]
for lg in logs:
    n = 1 << lg
    a, b, c = rand_fr(n), rand_fr(n), rand_fr(n)
    ms = C.c_float()
    fb.native.check(lib.fb_test_h(ctx.handle, lg, a.ctypes.data, b.ctypes.data, c.ctypes.data, None, C.byref(ms)))
    out[f"h_pipeline_ms_2^{lg}"] = ms.value
    print(lg, "H ms", ms.value, flush=True)
    k = rand_fr(n)
    for group, psz in ((1, 64), (2, 128)):
        if group == 2 and lg > 22:
            continue
        bases = np.zeros((n, psz), dtype=np.uint8)
        t = time.time()
        fb.native.check(lib.fb_test_fixed_base(ctx.handle, group, k.ctypes.data, n, bases.ctypes.data))
        out[f"fixed_base_g{group}_s_2^{lg}"] = time.time() - t
        res = np.zeros(psz, dtype=np.uint8)
        # plain per-window buckets, then the window-table mode the prover uses (tables built once, untimed)
        for mode, tag in ((0, "plain"), (1, "tables")):
            lib.fb_set_msm_tables(mode)
            fb.native.check(lib.fb_test_msm(ctx.handle, group, bases.ctypes.data, a.ctypes.data, n, res.ctypes.data, 3, C.byref(ms)))
            out[f"msm_g{group}_{tag}_ms_2^{lg}"] = ms.value
            out[f"msm_g{group}_{tag}_mpts_2^{lg}"] = n / ms.value / 1e3
            print(lg, "msm g", group, tag, ms.value, "ms", n / ms.value / 1e3, "Mpts/s", flush=True)
        lib.fb_set_msm_tables(-1)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
print(json.dumps(out))
