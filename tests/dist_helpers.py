"""Oracle-side computation of one rank's partial MSM sums (for the N>1 host-logic tests)."""
from oracle import bn254 as bn
from oracle import codec
from oracle import groth16 as og


def shard_partials(golden: dict, rank: int, world: int) -> bytes:
    """640-byte fb_prove_partial payload of `rank`, computed with the Python oracle using the
    same slicing rule as fb_pk_load_shard: indices [len*rank/world, len*(rank+1)/world)."""
    P = codec.bellman_params_parse(bytes.fromhex(golden["bellman_params_hex"]))
    gates = codec.parse_gates(bytes.fromhex(golden["gates_raw_hex"]))
    inp = [int(x, 16) for x in golden["inputs"]]
    aux = [int(x, 16) for x in golden["aux"]]
    asg = og.evaluate_r1cs(gates, inp, aux)
    h = og.h_coefficients(asg.a, asg.b, asg.c)
    m = len(h) + 1
    exp = m.bit_length() - 1
    # the library stores h bit-reversed: position p holds coefficient brev(p)
    h_pos = [og._bitrev(p, exp) for p in range(m - 1)]
    w = inp + aux
    a_sc = inp + og.select(aux, asg.a_aux_density)
    b_sc = og.select(inp, asg.b_in_density) + og.select(aux, asg.b_aux_density)

    def sl(n):
        return n * rank // world, n * (rank + 1) // world

    def part(o, bases, scalars):
        return bn.to_affine(o, og.msm(o, bases, scalars))

    lo, hi = sl(m - 1)
    ph = part(bn.OPS1, [P.h[h_pos[p]] for p in range(lo, hi)], [h[h_pos[p]] for p in range(lo, hi)])
    lo, hi = sl(len(aux))
    pl = part(bn.OPS1, P.l[lo:hi], aux[lo:hi])
    lo, hi = sl(len(P.a))
    pa = part(bn.OPS1, P.a[lo:hi], a_sc[lo:hi])
    lo, hi = sl(len(P.b_g1))
    pb1 = part(bn.OPS1, P.b_g1[lo:hi], b_sc[lo:hi])
    pb2 = part(bn.OPS2, P.b_g2[lo:hi], b_sc[lo:hi])
    out = b""
    for p in (ph, pl, pa, pb1):
        out += codec.g1_raw(p) + bytes(64)
    return out + codec.g2_raw(pb2)
