import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fawkes_crypto_b200 as fb
ctx = fb.Context(0); lib = fb.native.lib
d = C.c_double()
lib.fb_probe_imad(ctx.handle, C.byref(d)); print("imad.wide plain MAC/s %.3e" % d.value)
for which, name in ((0, "madc chain MAC/s"), (1, "mul_ptx /s"), (3, "mul29 /s"), (4, "sqr29+add+sub /s")):
    for thr, bps in ((128, 1), (128, 2), (128, 4), (256, 4), (256, 8), (512, 4)):
        lib.fb_probe_rate(ctx.handle, which, thr, bps, C.byref(d))
        print(f"{name:18s} threads={thr:4d} blocks/SM={bps}  warps/SMSP={thr*bps/128:5.1f}  {d.value:.3e}")
