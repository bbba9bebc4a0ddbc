"""Regenerate tests/golden/synth_proofs.json (CPU only, ~15 min on 8 cores for 2^24): for the benchmark circuits of
SURVEY.md section 8(d) -- 2^20 rows (configs[2]) and 2^24 rows (configs[3]) -- the C++ oracle generates circuit and
witness (oracle/cpu_setup.cpp, equal to oracle/synth.py), runs setup with the fixed trapdoor and proves with the
fixed r, s (oracle/cpu_prover.cpp).  The fixture pins sha256 of the bellman Parameters bytes and the 256 proof
bytes; bench.py checks every GPU proof (N = 1, 2, 4, 8) and every CPU-arm proof against it."""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cpu  # noqa: E402

if __name__ == "__main__":
    sizes = [int(x) for x in sys.argv[1:]] or [12, 16, 20, 24]
    path = os.path.join(ROOT, "tests", "golden", "synth_proofs.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    cfg = {20: 3, 24: 4}
    th = cpu.hw_threads()
    for lg in sizes:
        seed = 0xFA3CE50000 + cfg.get(lg, 100 + lg)
        c = cpu.Circuit.synthetic(1 << lg, seed)
        td = cpu.synth_trapdoor(seed)
        p, _ = cpu.setup(c, td, th)
        t = time.time()
        pr, _, st = cpu.prove_circuit(p, c, td[5], td[6], th)
        out[str(lg)] = {"seed": seed, "n_aux": c.n_aux, "nnz": c.nnz, "params_bytes": p.size,
                        "params_sha256": hashlib.sha256(p.array).hexdigest(),
                        "proof_raw_hex": pr.hex(), "proof_sha256": hashlib.sha256(pr).hexdigest()}
        print(lg, out[str(lg)]["proof_sha256"], f"prove {time.time() - t:.1f} s", flush=True)
        json.dump(out, open(path, "w"), indent=1)
        del p, c
