// Groth16 verification on the host (BN254 optimal-ate pairing).
//
// Replaces bellman::groth16::{prepare_verifying_key, verify_proof} as called from
// fawkes-crypto/src/backend/bellman_groth16/verifier.rs:75-81 (restated in SURVEY.md
// App. C.6):  IC = ic_0 + sum x_i ic_{i+1};  accept iff
// e(A,B) * e(-IC,gamma) * e(-C,delta) * e(-alpha,beta) == 1.
// Host code, as in the reference (three Miller loops are not a hot path).
// Fq12 = Fq2[w]/(w^6 - xi), xi = 9 + u; untwist (x,y) -> (x w^2, y w^3).
#include "../../include/fawkes_b200.h"

#include <cstring>

#include "internal.h"
#include "verify_consts.h"

namespace fb {
namespace {

struct F12 {
  Fq2 c[6];
};

Fq2 xi_mul(const Fq2& a) {  // (a0 + a1 u)(9 + u) = 9a0 - a1 + (a0 + 9a1) u
  Fq t0 = dbl(dbl(dbl(a.c0)));
  t0 = add(t0, a.c0);
  Fq t1 = dbl(dbl(dbl(a.c1)));
  t1 = add(t1, a.c1);
  return {sub(t0, a.c1), add(t1, a.c0)};
}

F12 f12_one() {
  F12 r;
  for (auto& x : r.c) x = Fq2::zero();
  r.c[0] = Fq2::one();
  return r;
}
bool f12_eq(const F12& a, const F12& b) {
  for (int i = 0; i < 6; i++)
    if (a.c[i] != b.c[i]) return false;
  return true;
}
F12 f12_mul(const F12& a, const F12& b) {
  Fq2 t[11];
  for (auto& x : t) x = Fq2::zero();
  for (int i = 0; i < 6; i++) {
    if (a.c[i].is_zero()) continue;
    for (int j = 0; j < 6; j++) {
      if (b.c[j].is_zero()) continue;
      t[i + j] = add(t[i + j], mul(a.c[i], b.c[j]));
    }
  }
  F12 r;
  for (int k = 0; k < 6; k++) r.c[k] = k + 6 < 11 ? add(t[k], xi_mul(t[k + 6])) : t[k];
  return r;
}
F12 f12_conj(const F12& a) {  // a^(p^6): w -> -w
  F12 r = a;
  for (int i = 1; i < 6; i += 2) r.c[i] = neg(a.c[i]);
  return r;
}
Fq2 f2_conj(const Fq2& a) { return {a.c0, neg(a.c1)}; }
Fq2 frob_w(int i) {
  Fq2 r;
  for (int j = 0; j < 8; j++) { r.c0.v[j] = FROB_W[i][0][j]; r.c1.v[j] = FROB_W[i][1][j]; }
  return {to_mont(r.c0), to_mont(r.c1)};
}
F12 f12_frob(const F12& a) {
  F12 r;
  for (int i = 0; i < 6; i++) r.c[i] = mul(f2_conj(a.c[i]), frob_w(i));
  return r;
}
F12 f12_inv(const F12& a) {
  // n = a * conj(a) lies in Fq2[v]/(v^3 - xi), v = w^2; invert there, then a^-1 = conj(a) / n
  F12 n = f12_mul(a, f12_conj(a));
  Fq2 c0 = n.c[0], c1 = n.c[2], c2 = n.c[4];
  Fq2 t0 = sub(sqr(c0), xi_mul(mul(c1, c2)));
  Fq2 t1 = sub(xi_mul(sqr(c2)), mul(c0, c1));
  Fq2 t2 = sub(sqr(c1), mul(c0, c2));
  Fq2 d = add(mul(c0, t0), xi_mul(add(mul(c2, t1), mul(c1, t2))));
  Fq2 di = inv(d);
  F12 ni;
  for (auto& x : ni.c) x = Fq2::zero();
  ni.c[0] = mul(t0, di);
  ni.c[2] = mul(t1, di);
  ni.c[4] = mul(t2, di);
  return f12_mul(f12_conj(a), ni);
}
F12 f12_pow(const F12& a, const uint32_t* e, int limbs) {
  F12 r = f12_one();
  bool started = false;
  for (int i = limbs * 32 - 1; i >= 0; i--) {
    if (started) r = f12_mul(r, r);
    if ((e[i >> 5] >> (i & 31)) & 1) {
      r = started ? f12_mul(r, a) : a;
      started = true;
    }
  }
  return r;
}

// line through twist points T, Q (tangent when equal) evaluated at P; T <- T + Q
F12 line(G2Affine& T, const G2Affine& Q, const G1Affine& P) {
  Fq2 lam;
  if (T.x == Q.x && T.y == Q.y) {
    Fq2 x2 = sqr(T.x);
    lam = mul(add(dbl(x2), x2), inv(dbl(T.y)));
  } else {
    lam = mul(sub(Q.y, T.y), inv(sub(Q.x, T.x)));
  }
  Fq2 x3 = sub(sub(sqr(lam), T.x), Q.x);
  Fq2 y3 = sub(mul(lam, sub(T.x, x3)), T.y);
  F12 l;
  for (auto& x : l.c) x = Fq2::zero();
  l.c[0] = {P.y, Fq::zero()};
  l.c[1] = neg(Fq2{mul(lam.c0, P.x), mul(lam.c1, P.x)});
  l.c[3] = sub(mul(lam, T.x), T.y);
  T = {x3, y3};
  return l;
}

G2Affine twist_frob(const G2Affine& q) {
  return {mul(f2_conj(q.x), frob_w(2)), mul(f2_conj(q.y), frob_w(3))};
}

F12 miller_loop(const G1Affine& P, const G2Affine& Q) {
  if (P.is_inf() || Q.is_inf()) return f12_one();
  F12 f = f12_one();
  G2Affine T = Q;
  for (int i = ATE_LOOP_BITS - 2; i >= 0; i--) {
    F12 l = line(T, T, P);
    f = f12_mul(f12_mul(f, f), l);
    if ((ATE_LOOP[i >> 5] >> (i & 31)) & 1) {
      l = line(T, Q, P);
      f = f12_mul(f, l);
    }
  }
  G2Affine Q1 = twist_frob(Q);
  G2Affine Q2 = twist_frob(Q1);
  G2Affine nQ2 = {Q2.x, neg(Q2.y)};
  f = f12_mul(f, line(T, Q1, P));
  f = f12_mul(f, line(T, nQ2, P));
  return f;
}

F12 final_exp(const F12& f) {
  F12 f1 = f12_mul(f12_conj(f), f12_inv(f));      // f^(p^6 - 1)
  F12 f2 = f12_mul(f12_frob(f12_frob(f1)), f1);   // ^(p^2 + 1)
  return f12_pow(f2, HARD_EXP, HARD_EXP_LIMBS);
}

}  // namespace
}  // namespace fb

using namespace fb;

extern "C" int fb_verify(const uint8_t* vk_raw, uint32_t n_ic, const uint8_t proof_raw[256],
                         const uint64_t* inputs, uint32_t n_inputs, int* ok) try {
  if (!vk_raw || !proof_raw || !ok || (!inputs && n_inputs)) { set_error("fb_verify: bad argument"); return FB_ERR_ARG; }
  *ok = 0;
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
  G1Affine ic0;
  memcpy(&ic0, ic, 64);
  G1XYZZ acc = G1XYZZ::from_affine(ic0);
  for (uint32_t i = 0; i < n_inputs; i++) {
    G1Affine p;
    memcpy(&p, ic + 64 * (i + 1), 64);
    Fr x;
    memcpy(x.v, inputs + 4 * i, 32);
    x = from_mont(x);
    acc = add_cold(acc, scalar_mul(G1XYZZ::from_affine(p), x.v));
  }
  G1Affine IC = to_affine(acc);
  F12 f = miller_loop(A, B);
  f = f12_mul(f, miller_loop(neg(IC), gamma));
  f = f12_mul(f, miller_loop(neg(C), delta));
  f = f12_mul(f, miller_loop(neg(alpha), beta));
  *ok = f12_eq(final_exp(f), f12_one()) ? 1 : 0;
  return FB_OK;
} FB_ABI_CATCH_INT
