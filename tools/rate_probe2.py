import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fawkes_crypto_b200 as fb
ctx = fb.Context(0); lib = fb.native.lib
d = C.c_double()
lib.fb_probe_imad(ctx.handle, C.byref(d)); print("8 acc, shared multiplier (fb_probe_imad) %.3e" % d.value)
for which, name in ((9, "8 acc, 3 distinct regs"), (10, "1 acc chain, distinct")):
    for thr, bps in ((128, 1), (128, 4), (256, 8)):
        lib.fb_probe_rate(ctx.handle, which, thr, bps, C.byref(d))
        print(f"{name:24s} warps/SMSP={thr*bps/128:5.1f}  {d.value:.3e} MAC/s")
