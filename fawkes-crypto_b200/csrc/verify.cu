// Groth16 verification on the host (BN254 optimal-ate pairing).
//
// Replaces bellman::groth16::{prepare_verifying_key, verify_proof} as called from
// fawkes-crypto/src/backend/bellman_groth16/verifier.rs:75-81 (restated in SURVEY.md
// App. C.6):  IC = ic_0 + sum x_i ic_{i+1};  accept iff
// e(A,B) * e(-IC,gamma) * e(-C,delta) * e(-alpha,beta) == 1.
// Host code, as in the reference (three Miller loops are not a hot path).
// Fq12 = Fq2[w]/(w^6 - xi), xi = 9 + u; untwist (x,y) -> (x w^2, y w^3).
#include "../../include/fawkes_b200.h"

#include <chrono>
#include <cstdlib>
#include <cstring>

#include "internal.h"
#include "verify_consts.h"

namespace fb {
namespace {

// All arithmetic below runs on the 64-bit-limb host types of host_fq.h (HFq, HFq2); the tower is
// Fq6 = Fq2[v]/(v^3 - xi), Fq12 = Fq6[w]/(w^2 - v), i.e. sum c_i w^i with c = (x.a, y.a, x.b, y.b, x.c, y.c).
typedef HFq F1;
typedef HFq2 F2;

inline F2 f2_conj(const F2& a) { return {a.c0, neg(a.c1)}; }
inline F2 f2_scale(const F2& a, const F1& k) { return {mul(a.c0, k), mul(a.c1, k)}; }
inline F2 xi_mul(const F2& a) {  // (a0 + a1 u)(9 + u) = 9a0 - a1 + (a0 + 9a1) u
  F1 t0 = dbl(dbl(dbl(a.c0)));
  t0 = add(t0, a.c0);
  F1 t1 = dbl(dbl(dbl(a.c1)));
  t1 = add(t1, a.c1);
  return {sub(t0, a.c1), add(t1, a.c0)};
}

struct F6 {
  F2 a, b, c;  // a + b v + c v^2
};
inline F6 f6_zero() { return {F2::zero(), F2::zero(), F2::zero()}; }
inline F6 f6_add(const F6& x, const F6& y) { return {add(x.a, y.a), add(x.b, y.b), add(x.c, y.c)}; }
inline F6 f6_sub(const F6& x, const F6& y) { return {sub(x.a, y.a), sub(x.b, y.b), sub(x.c, y.c)}; }
inline F6 f6_neg(const F6& x) { return {neg(x.a), neg(x.b), neg(x.c)}; }
inline F6 f6_mul_v(const F6& x) { return {xi_mul(x.c), x.a, x.b}; }
inline F6 f6_scale(const F6& x, const F1& k) { return {f2_scale(x.a, k), f2_scale(x.b, k), f2_scale(x.c, k)}; }
F6 f6_mul(const F6& x, const F6& y) {  // Karatsuba: 6 Fq2 products
  const F2 v0 = mul(x.a, y.a), v1 = mul(x.b, y.b), v2 = mul(x.c, y.c);
  const F2 t0 = sub(sub(mul(add(x.b, x.c), add(y.b, y.c)), v1), v2);
  const F2 t1 = sub(sub(mul(add(x.a, x.b), add(y.a, y.b)), v0), v1);
  const F2 t2 = sub(sub(mul(add(x.a, x.c), add(y.a, y.c)), v0), v2);
  return {add(v0, xi_mul(t0)), add(t1, xi_mul(v2)), add(t2, v1)};
}
F6 f6_mul_sparse(const F6& x, const F2& l1, const F2& l3) {  // x * (l1 + l3 v)
  return {add(mul(x.a, l1), xi_mul(mul(x.c, l3))), add(mul(x.a, l3), mul(x.b, l1)), add(mul(x.b, l3), mul(x.c, l1))};
}
F6 f6_inv(const F6& x) {
  const F2 t0 = sub(sqr(x.a), xi_mul(mul(x.b, x.c)));
  const F2 t1 = sub(xi_mul(sqr(x.c)), mul(x.a, x.b));
  const F2 t2 = sub(sqr(x.b), mul(x.a, x.c));
  const F2 d = add(mul(x.a, t0), xi_mul(add(mul(x.c, t1), mul(x.b, t2))));
  const F2 di = inv(d);
  return {mul(t0, di), mul(t1, di), mul(t2, di)};
}

struct F12 {
  F6 x, y;  // x + y w
};
F12 f12_one() { return {{F2::one(), F2::zero(), F2::zero()}, f6_zero()}; }
bool f12_eq(const F12& a, const F12& b) {
  return a.x.a == b.x.a && a.x.b == b.x.b && a.x.c == b.x.c && a.y.a == b.y.a && a.y.b == b.y.b && a.y.c == b.y.c;
}
F12 f12_mul(const F12& a, const F12& b) {  // 3 Fq6 products
  const F6 aa = f6_mul(a.x, b.x), bb = f6_mul(a.y, b.y);
  const F6 m = f6_mul(f6_add(a.x, a.y), f6_add(b.x, b.y));
  return {f6_add(aa, f6_mul_v(bb)), f6_sub(f6_sub(m, aa), bb)};
}
F12 f12_sqr(const F12& a) {  // complex squaring: 2 Fq6 products
  const F6 ab = f6_mul(a.x, a.y);
  const F6 t = f6_mul(f6_add(a.x, a.y), f6_add(a.x, f6_mul_v(a.y)));
  return {f6_sub(f6_sub(t, ab), f6_mul_v(ab)), f6_add(ab, ab)};
}
// f * (l0 + l1 w + l3 w^3), l0 in Fq: the value of a line at a G1 point (untwist (x, y) -> (x w^2, y w^3))
F12 f12_mul_line(const F12& f, const F1& l0, const F2& l1, const F2& l3) {
  const F6 t = f6_mul_sparse(f.y, l1, l3), u = f6_mul_sparse(f.x, l1, l3);
  return {f6_add(f6_scale(f.x, l0), f6_mul_v(t)), f6_add(u, f6_scale(f.y, l0))};
}
F12 f12_conj(const F12& a) { return {a.x, f6_neg(a.y)}; }  // a^(p^6): w -> -w
F12 f12_inv(const F12& a) {
  const F6 t = f6_inv(f6_sub(f6_mul(a.x, a.x), f6_mul_v(f6_mul(a.y, a.y))));
  return {f6_mul(a.x, t), f6_neg(f6_mul(a.y, t))};
}
const F2* frob_consts() {  // xi^(i (p-1)/6), i = 0..5, Montgomery
  static const struct Table {
    F2 g[6];
    Table() {
      for (int i = 0; i < 6; i++) {
        Fq2 r;
        for (int j = 0; j < 8; j++) { r.c0.v[j] = FROB_W[i][0][j]; r.c1.v[j] = FROB_W[i][1][j]; }
        g[i] = F2::from(Fq2{to_mont(r.c0), to_mont(r.c1)});
      }
    }
  } table;
  return table.g;
}
F12 f12_frob(const F12& a) {
  const F2* g = frob_consts();
  return {{mul(f2_conj(a.x.a), g[0]), mul(f2_conj(a.x.b), g[2]), mul(f2_conj(a.x.c), g[4])},
          {mul(f2_conj(a.y.a), g[1]), mul(f2_conj(a.y.b), g[3]), mul(f2_conj(a.y.c), g[5])}};
}
F12 f12_pow(const F12& a, const uint32_t* e, int limbs) {
  F12 r = f12_one();
  bool started = false;
  for (int i = limbs * 32 - 1; i >= 0; i--) {
    if (started) r = f12_sqr(r);
    if ((e[i >> 5] >> (i & 31)) & 1) {
      r = started ? f12_mul(r, a) : a;
      started = true;
    }
  }
  return r;
}

// 1/x for n Fq2 values with ONE field inversion (Montgomery's trick on the norms); 1/0 := 0, as a^(p-2) gives
void f2_batch_inv(F2* x, int n) {
  F1 norm[8], pre[8];
  F1 acc = F1::one();
  for (int k = 0; k < n; k++) {
    norm[k] = add(sqr(x[k].c0), sqr(x[k].c1));
    pre[k] = acc;
    if (!norm[k].is_zero()) acc = mul(acc, norm[k]);
  }
  acc = inv(acc);
  for (int k = n - 1; k >= 0; k--) {
    if (norm[k].is_zero()) { x[k] = F2::zero(); continue; }
    const F1 ni = mul(acc, pre[k]);
    acc = mul(acc, norm[k]);
    x[k] = {mul(x[k].c0, ni), neg(mul(x[k].c1, ni))};
  }
}

struct MillerPair {
  F1 px, py;   // G1 point
  F2 qx, qy;   // G2 point (on the twist)
  F2 tx, ty;   // running multiple of Q
};
// One chord / tangent step of every pair: T <- T + S (S = T for a tangent), f <- f * line_{T,S}(P).
// sx/sy: the second point per pair (ignored where tangent[k]).
void line_step(F12& f, MillerPair* ps, int n, const F2* sx, const F2* sy, const bool* tangent) {
  F2 den[8];
  for (int k = 0; k < n; k++) den[k] = tangent[k] ? dbl(ps[k].ty) : sub(sx[k], ps[k].tx);
  f2_batch_inv(den, n);
  for (int k = 0; k < n; k++) {
    MillerPair& p = ps[k];
    F2 lam, ox;
    if (tangent[k]) {
      const F2 x2 = sqr(p.tx);
      lam = mul(add(dbl(x2), x2), den[k]);
      ox = p.tx;
    } else {
      lam = mul(sub(sy[k], p.ty), den[k]);
      ox = sx[k];
    }
    const F2 x3 = sub(sub(sqr(lam), p.tx), ox);
    const F2 y3 = sub(mul(lam, sub(p.tx, x3)), p.ty);
    f = f12_mul_line(f, p.py, neg(f2_scale(lam, p.px)), sub(mul(lam, p.tx), p.ty));
    p.tx = x3;
    p.ty = y3;
  }
}

// prod_k e(P_k, Q_k) before the final exponentiation: one shared squaring of f per loop bit
F12 multi_miller_loop(MillerPair* ps, int n) {
  F12 f = f12_one();
  if (n == 0) return f;
  const F2* g = frob_consts();
  bool tan[8], chord[8];
  F2 sx[8], sy[8];
  for (int k = 0; k < n; k++) { ps[k].tx = ps[k].qx; ps[k].ty = ps[k].qy; tan[k] = true; }
  for (int i = ATE_LOOP_BITS - 2; i >= 0; i--) {
    f = f12_sqr(f);
    line_step(f, ps, n, nullptr, nullptr, tan);
    if ((ATE_LOOP[i >> 5] >> (i & 31)) & 1) {
      for (int k = 0; k < n; k++) {
        sx[k] = ps[k].qx;
        sy[k] = ps[k].qy;
        chord[k] = ps[k].tx == ps[k].qx && ps[k].ty == ps[k].qy;  // T == Q: the chord is the tangent
      }
      line_step(f, ps, n, sx, sy, chord);
    }
  }
  // Q1 = pi(Q), Q2 = pi^2(Q): f *= line(T, Q1), then line(T + Q1, -Q2)
  F2 q2x[8], q2y[8];
  for (int k = 0; k < n; k++) {
    sx[k] = mul(f2_conj(ps[k].qx), g[2]);
    sy[k] = mul(f2_conj(ps[k].qy), g[3]);
    q2x[k] = mul(f2_conj(sx[k]), g[2]);
    q2y[k] = neg(mul(f2_conj(sy[k]), g[3]));
    chord[k] = ps[k].tx == sx[k] && ps[k].ty == sy[k];
  }
  line_step(f, ps, n, sx, sy, chord);
  for (int k = 0; k < n; k++) chord[k] = ps[k].tx == q2x[k] && ps[k].ty == q2y[k];
  line_step(f, ps, n, q2x, q2y, chord);
  return f;
}

// f^x for the BN parameter x = 4965661367192848881 (63 bits, weight 28)
F12 f12_pow_x(const F12& a) {
  const uint64_t x = 4965661367192848881ull;
  F12 r = a;
  for (int i = 61; i >= 0; i--) {
    r = f12_sqr(r);
    if ((x >> i) & 1) r = f12_mul(r, a);
  }
  return r;
}

// f^((p^12 - 1)/r).  Easy part (p^6 - 1)(p^2 + 1), then the hard part (p^4 - p^2 + 1)/r =
// p^3 + (6x^2 + 1) p^2 + (-36x^3 - 18x^2 - 12x + 1) p + (-36x^3 - 30x^2 - 18x - 2) evaluated with three
// exponentiations by x, Frobenius maps and the vectorial addition chain of Scott et al. (y0 y1^2 y2^6 y3^12 y4^18
// y5^30 y6^36); after the easy part the element is unitary, so an inverse is a conjugation.  The chain gives
// EXACTLY that exponent (integer identity, re-checked in tests/test_abi.py; fb_test_pairing compares the value with
// the plain 762-bit square-and-multiply over HARD_EXP, kept below for that purpose).
F12 final_exp(const F12& f) {
  const F12 f1 = f12_mul(f12_conj(f), f12_inv(f));      // f^(p^6 - 1)
  const F12 m = f12_mul(f12_frob(f12_frob(f1)), f1);    // ^(p^2 + 1)
  const F12 mx = f12_pow_x(m), mx2 = f12_pow_x(mx), mx3 = f12_pow_x(mx2);
  const F12 mp = f12_frob(m), mp2 = f12_frob(mp), mp3 = f12_frob(mp2);
  const F12 y0 = f12_mul(f12_mul(mp, mp2), mp3);
  const F12 y1 = f12_conj(m);
  const F12 y2 = f12_frob(f12_frob(mx2));
  const F12 y3 = f12_conj(f12_frob(mx));
  const F12 y4 = f12_conj(f12_mul(mx, f12_frob(mx2)));
  const F12 y5 = f12_conj(mx2);
  const F12 y6 = f12_conj(f12_mul(mx3, f12_frob(mx3)));
  F12 t0 = f12_mul(f12_mul(f12_sqr(y6), y4), y5);
  F12 t1 = f12_mul(f12_mul(y3, y5), t0);
  t0 = f12_mul(t0, y2);
  t1 = f12_mul(f12_sqr(t1), t0);
  t1 = f12_sqr(t1);
  t0 = f12_mul(t1, y1);
  t1 = f12_mul(t1, y0);
  t0 = f12_sqr(t0);
  return f12_mul(t0, t1);
}
F12 final_exp_plain(const F12& f) {  // the same value by plain square-and-multiply (self-check only)
  const F12 f1 = f12_mul(f12_conj(f), f12_inv(f));
  const F12 m = f12_mul(f12_frob(f12_frob(f1)), f1);
  return f12_pow(m, HARD_EXP, HARD_EXP_LIMBS);
}

}  // namespace
}  // namespace fb

namespace fb {
G1Affine g1_generator();  // setup.cu
G2Affine g2_generator();
}  // namespace fb

using namespace fb;

// Self-check of the host pairing (no device needed).  For `n` rounds with a splitmix64 stream from `seed`:
//  (0) HFq inv (binary Euclid) == a^(p-2) on 64 values per round;
//  (1) the addition-chain final exponentiation equals plain square-and-multiply over (p^4 - p^2 + 1)/r on a random
//      element of Fq12;  (2) bilinearity: e(a G1, b G2) == e(G1, G2)^(a b) for random 62-bit a, b, and e(G1, G2) != 1;
//  (3) the multi-pair loop: e(P, Q) * e(-P, Q) == 1.   *failures = number of failed checks.
extern "C" int fb_test_pairing(uint64_t seed, int n, int* failures) try {
  if (!failures || n < 0) return FB_ERR_ARG;
  uint64_t st = seed;
  auto next = [&]() {
    uint64_t z = (st += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  };
  auto rnd_f1 = [&]() {
    HFq x;
    for (auto& l : x.v) l = next();
    x.v[3] &= (1ull << 60) - 1;  // < p
    return x;
  };
  auto rnd_f2 = [&]() { HFq2 x = {rnd_f1(), rnd_f1()}; return x; };
  typedef XYZZ<HFq> H1;
  typedef XYZZ<HFq2> H2;
  const G1Affine g1 = g1_generator();
  const G2Affine g2 = g2_generator();
  const H1 G1h = {HFq::from(g1.x), HFq::from(g1.y), HFq::one(), HFq::one()};
  const H2 G2h = {HFq2::from(g2.x), HFq2::from(g2.y), HFq2::one(), HFq2::one()};
  auto pairing = [&](const Affine<HFq>& P, const Affine<HFq2>& Q) {
    MillerPair mp;
    mp.px = P.x; mp.py = P.y; mp.qx = Q.x; mp.qy = Q.y;
    return final_exp(multi_miller_loop(&mp, 1));
  };
  const F12 e0 = pairing(to_affine(G1h), to_affine(G2h));
  int bad = 0;
  if (f12_eq(e0, f12_one())) bad++;
  for (int i = 0; i < n; i++) {
    for (int k = 0; k < 64; k++) {  // (0) the binary-Euclid inverse against a^(p-2), edge values included
      HFq a = rnd_f1();
      if (i == 0 && k == 0) a = HFq::zero();
      if (i == 0 && k == 1) a = HFq::one();
      if (i == 0 && k == 2) a = neg(HFq::one());
      if (i == 0 && k == 3) { a = HFq::zero(); a.v[0] = 1; }
      if (i == 0 && k == 4) { a = HFq::zero(); a.v[0] = 2; }
      if (inv(a) != inv_fermat(a)) bad++;
    }
    const F12 f = {{rnd_f2(), rnd_f2(), rnd_f2()}, {rnd_f2(), rnd_f2(), rnd_f2()}};
    if (!f12_eq(final_exp(f), final_exp_plain(f))) bad++;
    const uint64_t a = next() >> 2, b = next() >> 2;
    const uint32_t al[8] = {(uint32_t)a, (uint32_t)(a >> 32), 0, 0, 0, 0, 0, 0};
    const uint32_t bl[8] = {(uint32_t)b, (uint32_t)(b >> 32), 0, 0, 0, 0, 0, 0};
    const Affine<HFq> P = to_affine(scalar_mul(G1h, al));
    const Affine<HFq2> Q = to_affine(scalar_mul(G2h, bl));
    const unsigned __int128 ab = (unsigned __int128)a * b;
    const uint32_t abl[4] = {(uint32_t)ab, (uint32_t)(ab >> 32), (uint32_t)(ab >> 64), (uint32_t)(ab >> 96)};
    if (!f12_eq(pairing(P, Q), f12_pow(e0, abl, 4))) bad++;
    MillerPair two[2];
    two[0].px = P.x; two[0].py = P.y; two[0].qx = Q.x; two[0].qy = Q.y;
    two[1] = two[0];
    two[1].py = neg(P.y);
    if (!f12_eq(final_exp(multi_miller_loop(two, 2)), f12_one())) bad++;
  }
  *failures = bad;
  return FB_OK;
} FB_ABI_CATCH_INT

extern "C" int fb_verify(const uint8_t* vk_raw, uint32_t n_ic, const uint8_t proof_raw[256],
                         const uint64_t* inputs, uint32_t n_inputs, int* ok) try {
  if (!vk_raw || !proof_raw || !ok || (!inputs && n_inputs)) { set_error("fb_verify: bad argument"); return FB_ERR_ARG; }
  *ok = 0;
  auto tq00 = std::chrono::steady_clock::now();
  if (n_inputs + 1 != n_ic) {
    set_error("MalformedVerifyingKey: %u inputs for %u ic points", n_inputs, n_ic);
    return FB_ERR_VK;
  }
  // The reference converts every point with `from_raw_uncompressed_le(..).unwrap()` (group.rs:53-65,87-105) and
  // every input with `from_raw_repr(..).unwrap()` (mod.rs:105-120): limbs >= the modulus or a point off its curve
  // panic there.  Here they are a format error, so that no second bit pattern of the same residue (x and x + r)
  // can verify for one statement.
  auto canonical = [](const uint8_t* p, int n_fq) {
    for (int i = 0; i < n_fq; i++) {
      uint32_t v[8];
      memcpy(v, p + 32 * i, 32);
      if (geq_mod<FqCfg>(v)) return false;
    }
    return true;
  };
  if (!canonical(vk_raw, 2 + 4 + 4 + 4) || !canonical(vk_raw + 448, 2 * (int)n_ic)) {
    set_error("verifying key holds a coordinate that is not a reduced field element");
    return FB_ERR_FORMAT;
  }
  if (!canonical(proof_raw, 8)) {
    set_error("proof holds a coordinate that is not a reduced field element");
    return FB_ERR_FORMAT;
  }
  for (uint32_t i = 0; i < n_inputs; i++) {
    uint32_t v[8];
    memcpy(v, inputs + 4 * (size_t)i, 32);
    if (geq_mod<FrCfg>(v)) {
      set_error("public input %u is not a reduced field element", i);
      return FB_ERR_FORMAT;
    }
  }
  G1Affine alpha, A, C;
  G2Affine beta, gamma, delta, B;
  memcpy(&alpha, vk_raw, 64);
  memcpy(&beta, vk_raw + 64, 128);
  memcpy(&gamma, vk_raw + 192, 128);
  memcpy(&delta, vk_raw + 320, 128);
  const uint8_t* ic = vk_raw + 448;
  memcpy(&A, proof_raw, 64);
  memcpy(&B, proof_raw + 64, 128);
  memcpy(&C, proof_raw + 192, 64);
  if (!on_curve(alpha, g1_b()) || !on_curve(beta, g2_b()) || !on_curve(gamma, g2_b()) || !on_curve(delta, g2_b())) {
    set_error("verifying-key point not on its curve");
    return FB_ERR_FORMAT;
  }
  for (uint32_t i = 0; i < n_ic; i++) {
    G1Affine p;
    memcpy(&p, ic + 64 * (size_t)i, 64);
    if (!on_curve(p, g1_b())) {
      set_error("verifying-key point ic[%u] not on the curve", i);
      return FB_ERR_FORMAT;
    }
  }
  if (!on_curve(A, g1_b()) || !on_curve(C, g1_b()) || !on_curve(B, g2_b())) {
    set_error("proof point not on its curve");
    return FB_ERR_FORMAT;
  }
  typedef XYZZ<HFq> H1;
  auto h1 = [](const G1Affine& p) { return p.is_inf() ? H1::inf() : H1{HFq::from(p.x), HFq::from(p.y), HFq::one(), HFq::one()}; };
  G1Affine ic0;
  memcpy(&ic0, ic, 64);
  H1 acc = h1(ic0);
  for (uint32_t i = 0; i < n_inputs; i++) {
    G1Affine p;
    memcpy(&p, ic + 64 * (i + 1), 64);
    Fr x;
    memcpy(x.v, inputs + 4 * i, 32);
    x = from_mont(x);
    acc = add(acc, scalar_mul(h1(p), x.v));
  }
  const Affine<HFq> IC = to_affine(acc);
  // e(A, B) * e(-IC, gamma) * e(-C, delta) * e(-alpha, beta): pairs with a point at infinity contribute 1
  MillerPair ps[4];
  int np = 0;
  auto push = [&](const HFq& px, const HFq& py, bool p_inf, const G2Affine& q) {
    if (p_inf || q.is_inf()) return;
    ps[np].px = px;
    ps[np].py = py;
    ps[np].qx = HFq2::from(q.x);
    ps[np].qy = HFq2::from(q.y);
    np++;
  };
  push(HFq::from(A.x), HFq::from(A.y), A.is_inf(), B);
  push(IC.x, neg(IC.y), IC.is_inf(), gamma);
  push(HFq::from(C.x), neg(HFq::from(C.y)), C.is_inf(), delta);
  push(HFq::from(alpha.x), neg(HFq::from(alpha.y)), alpha.is_inf(), beta);
  auto tq0 = std::chrono::steady_clock::now();
  const F12 f = multi_miller_loop(ps, np);
  auto tq1 = std::chrono::steady_clock::now();
  *ok = f12_eq(final_exp(f), f12_one()) ? 1 : 0;
  if (getenv("FB_TRACE")) {
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fprintf(stderr, "[fb trace] verify: checks + IC %.3f  miller %.3f  final exp %.3f ms\n", ms(tq00, tq0), ms(tq0, tq1),
            ms(tq1, std::chrono::steady_clock::now()));
  }
  return FB_OK;
} FB_ABI_CATCH_INT
