"""C-ABI surface and host-side logic, no GPU needed: the library loads, exports every symbol
include/fawkes_b200.h declares, and its host code (gate parsing, synthetic generator, framing,
pairing verifier, partial-sum combine) agrees with the oracle."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import codec
from oracle import groth16 as og
from oracle import synth
from tests.util import fr_np, fr_list

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import fawkes_crypto_b200 as fb
    hdr = open(os.path.join(ROOT, "include", "fawkes_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    for n in sorted(names):
        assert hasattr(fb.native.lib, n), f"{n} declared in the header but not exported"
    assert names == set(fb.native.SIGNATURES), names ^ set(fb.native.SIGNATURES)


def _split_params(arglist: str):
    arglist = arglist.strip()
    if arglist in ("", "void"):
        return []
    return [a.strip() for a in arglist.split(",")]


def test_rust_sys_crate_follows_the_header():
    """rust/fawkes-b200-sys cannot be compiled in this image (no Rust toolchain), so its extern block is checked
    against include/fawkes_b200.h textually: every bound function is declared in the header with the same number
    of arguments and the same return kind, the error codes and load flags carry the header's values, and every
    product entry point (everything above the benchmark / self-test helpers) is bound."""
    hdr = open(os.path.join(ROOT, "include", "fawkes_b200.h")).read()
    hdr_nc = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    decl = {}
    for ret, name, args in re.findall(r"^\s*([A-Za-z_][\w \*]*?)\s*\b(fb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr_nc, flags=re.M):
        decl[name] = (ret.strip(), _split_params(" ".join(args.split())))
    rs = open(os.path.join(ROOT, "rust", "fawkes-b200-sys", "src", "lib.rs")).read()
    rs_nc = re.sub(r"//.*", "", rs)
    bound = {}
    for name, args, ret in re.findall(r"pub fn (fb_[a-z0-9_]+)\s*\(([^)]*)\)\s*(->\s*[^;]+)?;", rs_nc):
        bound[name] = (_split_params(" ".join(args.split())), ret.replace("->", "").strip())
    assert len(bound) >= 30
    for name, (args, ret) in bound.items():
        assert name in decl, f"{name} is bound in Rust but not declared in the header"
        cret, cargs = decl[name]
        assert len(args) == len(cargs), f"{name}: {len(args)} arguments in Rust, {len(cargs)} in the header"
        want = {"int": "c_int", "void": "", "const char*": "*const c_char", "uint64_t": "u64"}[cret.replace(" *", "*")]
        assert ret == want, f"{name}: returns {ret!r} in Rust, {cret!r} in the header"
        for ra, ca in zip(args, cargs):
            is_ptr_c = "*" in ca or "[" in ca
            is_ptr_rs = ra.split(":", 1)[1].strip().startswith("*")
            assert is_ptr_c == is_ptr_rs, f"{name}: argument {ra!r} against {ca!r}"
    # constants
    for cname, val in re.findall(r"#define\s+(FB_(?:OK|ERR_[A-Z]+|LOAD_[A-Z_]+))\s+\(?(-?\d+)\)?", hdr):
        m = re.search(rf"pub const {cname}: c_int = (-?\d+);", rs)
        assert m and int(m.group(1)) == int(val), f"{cname} differs between the header and the Rust crate"
    # the product surface: everything declared before the synthetic-circuit / instrumentation helpers
    product = [n for n in decl if hdr_nc.index(n + "(") < hdr_nc.index("fb_circuit_synth(")]
    missing = sorted(set(product) - set(bound) - {"fb_circuit_from_raw_gates_gpu"})
    assert not missing, f"product entry points without a Rust binding: {missing}"
    # fb_pk_info: same fields in the same order
    c_fields = []
    for ty, names in re.findall(r"\b(uint32_t|uint64_t)\s+([\w, ]+);",
                                hdr_nc[hdr_nc.index("typedef struct fb_pk_info"):hdr_nc.index("} fb_pk_info;")]):
        c_fields += [(n.strip(), ty[4:6]) for n in names.split(",")]
    r_fields = re.findall(r"pub (\w+): u(32|64),", rs[rs.index("pub struct fb_pk_info"):rs.index('extern "C"')])
    assert c_fields == r_fields, (c_fields, r_fields)


def test_host_exceptions_stop_at_the_abi():
    """A std::bad_alloc inside the library (here: a 2^27-row synthetic circuit under a 6 GiB address-space limit) must
    come back as FB_ERR_HOST with a message, not unwind through the C ABI (abort / undefined behaviour in a Rust
    caller).  Runs in a child process because of the rlimit."""
    import subprocess
    import sys
    code = (
        "import resource, ctypes, sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import fawkes_crypto_b200 as fb\n"
        "resource.setrlimit(resource.RLIMIT_AS, (6 << 30, 6 << 30))\n"
        "out = ctypes.c_void_p()\n"
        "rc = fb.native.lib.fb_circuit_synth(1 << 27, 1, ctypes.byref(out))\n"
        "print(rc, fb.native.lib.fb_last_error().decode())\n"
    )
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    rc, msg = r.stdout.strip().split(" ", 1)
    assert int(rc) == -8 and "host exception" in msg and "bad_alloc" in msg, r.stdout


def test_host_verify_accepts_the_committed_golden_proofs():
    """fb_verify needs no device, so the product's pairing is checked here against keys made by the CPU chain
    (oracle/cpu_setup.cpp) and the committed golden proofs: the 2^12-row synthetic circuit and the two real circuits
    (configs[0] Poseidon Merkle proof, configs[1] EdDSA-Poseidon, eight distinct signatures).  Three different
    gamma / delta / ic sets; every proof is accepted, every proof with another statement's inputs is rejected."""
    import hashlib
    import bench
    import fawkes_crypto_b200 as fb
    from oracle import cpu

    def vk_of(pbuf, n_in):
        params = fb.Parameters(pbuf.array, 0, b"")
        assert params.n_in == n_in
        return params.get_vk()

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "synth_proofs.json")))["12"]
    circ = cpu.Circuit.synthetic(1 << 12, g["seed"])
    pbuf, _ = cpu.setup(circ, cpu.synth_trapdoor(g["seed"]), 4)
    assert hashlib.sha256(pbuf.array).hexdigest() == g["params_sha256"]
    vk = vk_of(pbuf, circ.n_in)
    proof = fb.Proof.from_raw(bytes.fromhex(g["proof_raw_hex"]))
    inputs = np.array(circ.inputs[1:], dtype=np.uint64)
    assert fb.verify(vk, proof, inputs) is True
    assert fb.verify(vk, proof, fr_np([12345])) is False

    z = np.load(os.path.join(ROOT, "tests", "golden", "bench_cfgs.npz"))
    for cfg in ("cfg1", "cfg2"):
        n_gates, n_in, n_aux = (int(v) for v in z[f"{cfg}_shape"])
        c = fb.Circuit.from_gates_blob(z[f"{cfg}_gates_brotli"].tobytes(), n_gates, n_in, n_aux)
        rp, cl, cf = bench.expand_csr(fb, c)
        oc = cpu.Circuit.from_csr(n_gates, n_in, n_aux, rp, cl, cf)
        pbuf, _ = cpu.setup(oc, z[f"{cfg}_trapdoor_r_s"], 4)
        assert hashlib.sha256(pbuf.array).digest() == z[f"{cfg}_params_sha256"].tobytes()
        vk = vk_of(pbuf, n_in)
        if cfg == "cfg1":
            cases = [(z["cfg1_proof"].tobytes(), z["cfg1_inputs"][1:])]
        else:
            cases = [(z["cfg2_proofs_first8"][i].tobytes(), z["cfg2_inputs"][i][1:]) for i in range(8)]
        for i, (raw, ins) in enumerate(cases):
            assert fb.verify(vk, fb.Proof.from_raw(raw), np.ascontiguousarray(ins)) is True, (cfg, i)
        if len(cases) > 1:  # another signature's public input
            assert fb.verify(vk, fb.Proof.from_raw(cases[0][0]), np.ascontiguousarray(cases[1][1])) is False
        else:
            assert fb.verify(vk, fb.Proof.from_raw(cases[0][0]), fr_np([7])) is False


def test_host_pairing_self_check():
    """fb_verify's pairing runs on the host (verifier.rs:75-81 is host code in the reference too).  The library's
    self-check compares, on seeded random values: the binary-Euclid field inverse with a^(p-2); the addition-chain
    final exponentiation with plain square-and-multiply over (p^4 - p^2 + 1)/r; e(aG1, bG2) with e(G1, G2)^(ab); and
    the shared-squaring multi-pair loop on e(P, Q) e(-P, Q) = 1."""
    import fawkes_crypto_b200 as fb
    bad = C.c_int(-1)
    fb.native.check(fb.native.lib.fb_test_pairing(20261017, 6, C.byref(bad)))
    assert bad.value == 0


def test_final_exponentiation_chain_is_the_exact_exponent():
    """The exponent identity behind verify.cu: final_exp -- with conjugation = -1 and Frobenius = *p on exponents,
    the chain y0 y1^2 y2^6 y3^12 y4^18 y5^30 y6^36 is EXACTLY (p^4 - p^2 + 1)/r, and that is HARD_EXP."""
    p, r, x = bn.P, bn.R, 4965661367192848881
    assert p == 36 * x**4 + 36 * x**3 + 24 * x**2 + 6 * x + 1 and r == 36 * x**4 + 36 * x**3 + 18 * x**2 + 6 * x + 1
    e = (p**4 - p**2 + 1) // r
    src = open(os.path.join(ROOT, "fawkes-crypto_b200", "csrc", "verify_consts.h")).read()
    limbs = re.search(r"HARD_EXP\[24\] = \{([^}]*)\}", src).group(1)
    assert sum(int(v.strip().rstrip("u"), 16) << (32 * i) for i, v in enumerate(limbs.split(","))) == e
    y0 = p + p**2 + p**3
    y1, y2, y3, y4, y5, y6 = -1, x**2 * p**2, -(x * p), -(x + x**2 * p), -(x**2), -(x**3 + x**3 * p)
    t0 = 2 * y6 + y4 + y5
    t1 = y3 + y5 + t0
    t0 = t0 + y2
    t1 = 2 * (2 * t1 + t0)
    t0 = t1 + y1
    t1 = t1 + y0
    assert 2 * t0 + t1 == e


def test_verify_constants_are_regenerated_from_the_curve_parameters():
    """csrc/verify_consts.h (hard-part exponent, ate loop count, Frobenius constants) is exactly what
    tools/gen_verify_consts.py derives from the BN parameter x."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_verify_consts", os.path.join(ROOT, "tools", "gen_verify_consts.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.P == bn.P and mod.R == bn.R
    assert mod.header() == open(os.path.join(ROOT, "fawkes-crypto_b200", "csrc", "verify_consts.h")).read()


def test_no_cpu_fallback_without_device():
    import fawkes_crypto_b200 as fb
    if fb.native.lib.fb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(fb.native.FbError) as e:
        fb.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "fawkes-crypto_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "oracle/" not in src.replace("oracle/synth.py is the Python restatement", ""), f


def test_synthetic_generator_matches_python_restatement():
    import fawkes_crypto_b200 as fb
    for n_rows in (3, 10, 257):
        seed = synth.SEED_BASE + 50 + n_rows
        gates, inp, aux = synth.synth_circuit(n_rows, seed)
        c = fb.Circuit.synthetic(n_rows, seed)
        wi, wa = c.witness()
        assert fr_list(wi) == inp and fr_list(wa) == aux
        sh = c.shape()
        assert sh["n_gates"] == len(gates) and sh["nnz"] == sum(len(x) for g in gates for x in g)
        import bench
        rp, cl, cf = bench.expand_csr(fb, c)
        for m in range(3):
            terms = [(codec.fr_unraw(cf[m][i].tobytes()), int(cl[m][i])) for i in range(len(cl[m]))]
            want = [(cv, idx if tag == 0 else 2 + idx) for g in gates for cv, (tag, idx) in g[m]]
            assert terms == want
        td = np.zeros((7, 4), dtype=np.uint64)
        fb.native.check(fb.native.lib.fb_synth_trapdoor(seed, td.ctypes.data))
        t, r, s = synth.synth_trapdoor(seed)
        assert fr_list(td) == [t.alpha, t.beta, t.gamma, t.delta, t.tau, r, s]


def test_gate_stream_parsing_and_errors():
    import fawkes_crypto_b200 as fb
    gates, inp, aux = synth.synth_circuit(30, 7)
    raw = b"".join(codec.gate_borsh(g) for g in gates)
    blob = codec.brotli_compress(raw)
    c = fb.Circuit.from_gates_blob(blob, len(gates), 2, len(aux))
    assert c.shape() == {"n_in": 2, "n_aux": len(aux), "n_gates": len(gates), "nnz": sum(len(x) for g in gates for x in g)}
    # truncated stream -> fewer gates than announced -> FB_ERR_FORMAT
    with pytest.raises(fb.native.FbError) as e:
        fb.Circuit.from_raw_gates(raw[:-3], len(gates), 2, len(aux))
    assert e.value.code == -3
    # coefficient >= r is "Wrong raw integer": the stream ends there (cs.rs:215-223)
    bad = bytearray(raw)
    bad[4:36] = (bn.R + 5).to_bytes(32, "little")
    with pytest.raises(fb.native.FbError):
        fb.Circuit.from_raw_gates(bytes(bad), len(gates), 2, len(aux))
    # variable index out of range
    with pytest.raises(fb.native.FbError):
        fb.Circuit.from_raw_gates(raw, len(gates), 2, len(aux) - 5)
    # empty circuit
    c0 = fb.Circuit.from_raw_gates(b"", 0, 1, 0)
    assert c0.shape()["n_gates"] == 0


def test_host_parser_equals_python_parse_on_random_streams():
    """The host parser is the yardstick of the GPU ingest (tests/test_gpu_ingest.py): check it term by term
    against the Python restatement of the stream format (oracle codec, cs.rs:184-223) on random streams
    with repeated, +1, -1 and fresh coefficients and ragged LC lengths."""
    import bench
    import fawkes_crypto_b200 as fb
    from tests.util import random_gate_blob
    for n_gates, terms in ((1, (3, 2, 1)), (64, (5, 0, 4)), (300, (3, 3, 1))):
        raw = random_gate_blob(n_gates, 3, 40, seed=n_gates, terms=terms)
        ref = codec.parse_gates(raw)
        assert len(ref) == n_gates
        c = fb.Circuit.from_raw_gates(raw, n_gates, 3, 40)
        rp, cl, cf = bench.expand_csr(fb, c)
        for m in range(3):
            want_cols = [(idx if tag == 0 else 3 + idx) for g in ref for _, (tag, idx) in g[m]]
            want_coef = [cc for g in ref for cc, _ in g[m]]
            assert list(rp[m]) == [0] + list(np.cumsum([len(g[m]) for g in ref]))
            assert list(cl[m]) == want_cols
            assert fr_list(cf[m]) == want_coef


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "synth_rows40.json")))


def test_host_verify_and_framing_on_golden(golden):
    import fawkes_crypto_b200 as fb
    pb = bytes.fromhex(golden["bellman_params_hex"])
    raw = bytes.fromhex(golden["gates_raw_hex"])
    params = fb.Parameters(pb, 38, codec.brotli_compress(raw), [True, False])
    p2 = fb.Parameters.read(params.write())
    assert p2.bellman_bytes == pb and p2.num_gates == 38 and p2.const_tracker == [True, False]
    with pytest.raises(IOError):
        fb.Parameters.read(params.write()[:20])
    proof = fb.Proof.from_raw(bytes.fromhex(golden["proof_raw_hex"]))
    assert proof.serialize().hex() == golden["proof_borsh_hex"]
    assert fb.Proof.deserialize(proof.serialize()).to_raw() == proof.to_raw()
    vk = params.get_vk()
    assert fb.VK.deserialize(vk.serialize()).to_raw() == vk.to_raw()
    inputs = fr_np([int(golden["inputs"][1], 16)])
    assert fb.verify(vk, proof, inputs) is True
    assert fb.verify(vk, proof, fr_np([int(golden["inputs"][1], 16) ^ 1])) is False
    # a different, well-formed proof (C replaced by A: on the curve, wrong statement) is rejected ...
    swapped = proof.to_raw()[:192] + proof.to_raw()[:64]
    assert fb.verify(vk, fb.Proof.from_raw(swapped), inputs) is False
    # ... while malformed encodings are errors, as in the reference (`from_raw_uncompressed_le(..).unwrap()`,
    # group.rs:53-65; `from_raw_repr(..).unwrap()`, mod.rs:105-120): a point off its curve,
    tampered = bytearray(proof.to_raw())
    tampered[200] ^= 1
    with pytest.raises(fb.native.FbError) as e:
        fb.verify(vk, fb.Proof.from_raw(bytes(tampered)), inputs)
    assert e.value.code == -3 and "curve" in str(e.value)
    # limbs that are not a reduced field element: x + r is another bit pattern of the same residue and must not
    # verify for the same statement (input aliasing),
    x_mont = int.from_bytes(inputs[0].tobytes(), "little")
    alias = np.frombuffer((x_mont + bn.R).to_bytes(32, "little"), dtype=np.uint64).reshape(1, 4).copy()
    with pytest.raises(fb.native.FbError) as e:
        fb.verify(vk, proof, alias)
    assert e.value.code == -3 and "input" in str(e.value)
    # the same for a proof coordinate (A.x + p) and for a verifying-key point off the curve
    ax = int.from_bytes(proof.to_raw()[:32], "little") + bn.P
    if ax < 1 << 256:
        with pytest.raises(fb.native.FbError):
            fb.verify(vk, fb.Proof.from_raw(ax.to_bytes(32, "little") + proof.to_raw()[32:]), inputs)
    bad_vk = fb.VK.deserialize(vk.serialize())
    bad_vk.ic[1] = fb.G1Point(bad_vk.ic[1].raw[:32] + bad_vk.ic[0].raw[32:])
    with pytest.raises(fb.native.FbError) as e:
        fb.verify(bad_vk, proof, inputs)
    assert e.value.code == -3
    with pytest.raises(fb.native.FbError) as e:   # reference: MalformedVerifyingKey -> panic
        fb.verify(vk, proof, fr_np([1, 2]))
    assert e.value.code == -7


def test_msm_window_plans_of_the_benchmark_sizes():
    """MsmPlan::make (host logic): the windows DESIGN.md section 4.4 quotes for the benchmark sizes, with window tables.
    n: points of one MSM (a whole key, or one rank's shard of it)."""
    import fawkes_crypto_b200 as fb

    def plan(n, table=1):
        c, W, tl = C.c_int(), C.c_int(), C.c_int()
        fb.native.check(fb.native.lib.fb_test_msm_plan(n, table, C.byref(c), C.byref(W), C.byref(tl)))
        assert W.value * c.value >= 255 and (W.value - 1) * c.value < 255      # digits cover the scalar, no spare window
        return c.value, W.value

    assert plan(1 << 24) == (20, 13)            # configs[3] on one GPU
    assert plan(1 << 23) == (20, 13)            # ... its 2-GPU shards
    assert plan(1 << 22) == (20, 13)            # ... 4-GPU shards (c = 19 would pile 2^22 top digits on 128 counters)
    assert plan(1 << 21) == (17, 15)            # ... 8-GPU shards
    assert plan(1 << 20) == (17, 15)            # configs[2]
    assert plan(35695616 // 8) == (20, 13)      # configs[4] l/a/b shards on 8 GPUs
    assert plan(8191) == (10, 26) and plan(4125) == (10, 26)     # configs[0] / [1]: small circuits keep a 4-bit top digit
    for n in (1, 2, 3, 100, 5000, 1 << 16, (1 << 26) - 1):
        for table in (0, 1):
            c, W = plan(n, table)
            assert 4 <= c <= 22


def g2_point_outside_the_subgroup(seed=1):
    """A point of the twist curve y^2 = x^3 + 3/(9+u) that is NOT in the r-torsion subgroup (the cofactor of
    BN254's G2 is ~2^254, so the first curve point found by trying x values is outside)."""
    b2 = bn.OPS2.b

    def f2_sqrt(a):   # p = 3 mod 4 (Adj, Rodriguez-Henriquez alg. 9)
        a1 = bn.f2_pow(a, (bn.P - 3) // 4)
        alpha = bn.f2_mul(bn.f2_sqr(a1), a)
        a0 = bn.f2_mul(bn.f2_pow(alpha, bn.P), alpha)
        if a0 == (bn.P - 1, 0):
            return None
        x0 = bn.f2_mul(a1, a)
        if alpha == (bn.P - 1, 0):
            return bn.f2_mul((0, 1), x0)
        b = bn.f2_pow(bn.f2_add((1, 0), alpha), (bn.P - 1) // 2)
        return bn.f2_mul(b, x0)

    x = (seed, 1)
    while True:
        rhs = bn.f2_add(bn.f2_mul(bn.f2_sqr(x), x), b2)
        y = f2_sqrt(rhs)
        if y is not None and bn.f2_sqr(y) == rhs:
            pt = (x, y)
            assert bn.on_curve(bn.OPS2, pt)
            if bn.pt_mul(bn.OPS2, pt, bn.R) is not None:
                return pt
        x = (x[0] + 1, 1)


def test_verifying_key_points_are_read_checked(golden):
    """bellman's VerifyingKey::read decodes with `into_affine()` (on curve AND in the r-torsion subgroup), and the
    decoder itself rejects compression flags and dirty infinity encodings: the host decoder used for the
    verifying-key prefix (here through fb_prove_finish) does the same."""
    import fawkes_crypto_b200 as fb
    from tests.dist_helpers import shard_partials
    pb = bytearray(bytes.fromhex(golden["bellman_params_hex"]))
    parts = np.frombuffer(b"".join(shard_partials(golden, rank, 2) for rank in range(2)), dtype=np.uint8).copy()
    r, s = fr_np([int(golden["r"], 16)])[0], fr_np([int(golden["s"], 16)])[0]
    out = np.zeros(256, dtype=np.uint8)

    def finish(buf):
        b = bytes(buf)
        return fb.native.lib.fb_prove_finish(fb.native.ptr(b), len(b), parts.ctypes.data, 2, r.ctypes.data,
                                             s.ctypes.data, out.ctypes.data)

    assert finish(pb) == 0
    bad = bytearray(pb)
    bad[128:256] = codec.g2_uncompressed(g2_point_outside_the_subgroup())      # beta_g2: on the curve, wrong subgroup
    assert finish(bad) == -3
    bad = bytearray(pb)
    bad[0] |= 0x80                                                              # compression flag on alpha_g1
    assert finish(bad) == -3
    bad = bytearray(pb)
    bad[384:448] = bytes([0x40]) + bytes(62) + bytes([1])                       # delta_g1: dirty infinity encoding
    assert finish(bad) == -3
    bad = bytearray(pb)
    bad[64:128] = (bn.P).to_bytes(32, "big") + bad[96:128]                      # beta_g1.x = p: not a field element
    assert finish(bad) == -3


def test_prove_finish_combines_partials(golden):
    """fb_prove_finish (host): partial sums from two base shards -> the golden proof."""
    import fawkes_crypto_b200 as fb
    from tests.dist_helpers import shard_partials
    pb = bytes.fromhex(golden["bellman_params_hex"])
    parts = b"".join(shard_partials(golden, rank, 2) for rank in range(2))
    r, s = fr_np([int(golden["r"], 16)])[0], fr_np([int(golden["s"], 16)])[0]
    out = np.zeros(256, dtype=np.uint8)
    pbuf = np.frombuffer(parts, dtype=np.uint8).copy()
    fb.native.check(fb.native.lib.fb_prove_finish(fb.native.ptr(pb), len(pb), pbuf.ctypes.data, 2, r.ctypes.data,
                                                  s.ctypes.data, out.ctypes.data))
    assert out.tobytes().hex() == golden["proof_raw_hex"]
