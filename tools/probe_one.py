import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fawkes_crypto_b200 as fb
ctx = fb.Context(0); lib = fb.native.lib
d = C.c_double()
for which in [int(x) for x in os.environ.get("WHICH", "3,6,1").split(",")]:
    lib.fb_probe_rate(ctx.handle, which, 256, 8, C.byref(d)); print(which, "%.3e" % d.value)
