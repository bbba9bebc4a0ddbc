"""ncu target for the integer-pipe evidence (profiles/r02_imad_pipe.txt): runs the IMAD probe (k_probe_imad, the
kernel bench.py takes roofline.peak from) and one serial-schedule prove of 2^LOG rows, so that one ncu pass
    ncu --metrics sm__inst_executed_pipe_fmaheavy.sum,... -k regex:'k_probe_imad$|k_accumulate'
reports the fmaheavy-pipe instruction counts and active cycles of the probe and of k_accumulate side by side."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fawkes_crypto_b200 as fb
import bench

ctx = fb.Context(0)
v = C.c_double()
fb.native.check(fb.native.lib.fb_probe_imad(ctx.handle, C.byref(v)))
print("fb_probe_imad", v.value, flush=True)
lg = int(os.environ.get("LOG", "20"))
circ, params, tdi, _ = bench.make_case(fb, ctx, lg)
fb.native.lib.fb_set_serial(1)
wi, wa = circ.witness()
for i in range(int(os.environ.get("REPS", "1"))):
    inputs, proof = fb.groth16.prove_with_rs(params, wi, wa, tdi[5], tdi[6], ctx)
print(bench.sha_check(proof.to_raw(), lg))
