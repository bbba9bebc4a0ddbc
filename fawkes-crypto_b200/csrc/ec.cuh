// BN254 G1 (over Fq) and G2 (over Fq2) group arithmetic, a = 0 short Weierstrass.
//
// Replaces the curve arithmetic of the un-vendored pairing_ce bn256 crate that the
// reference reaches through `bellman::pairing::CurveAffine` (call sites
// fawkes-crypto/src/backend/bellman_groth16/group.rs:53-123).  Affine points use the
// reference's in-memory convention: raw Montgomery limbs, infinity <=> all coordinates
// zero (group.rs:55,71-72,89-93,111-112).  (0,0) is not on either curve (b != 0), so the
// encoding is unambiguous.
//
// Accumulators use extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ,
// ZZ^3 = ZZZ^2): mixed add 8M+2S, no inversion; infinity <=> ZZ == 0.
#pragma once
#include "ff.cuh"

namespace fb {

template <class F>
struct Affine {
  F x, y;
  FB_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
  FB_HD static Affine inf() { return {F::zero(), F::zero()}; }
};

template <class F>
struct XYZZ {
  F x, y, zz, zzz;
  FB_HD bool is_inf() const { return zz.is_zero(); }
  FB_HD static XYZZ inf() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
  FB_HD static XYZZ from_affine(const Affine<F>& p) {
    if (p.is_inf()) return inf();
    return {p.x, p.y, F::one(), F::one()};
  }
};

// 2*P for affine P (mdbl-2008-s-1)
template <class F>
FB_HD XYZZ<F> dbl_affine(const Affine<F>& p) {
  if (p.is_inf() || p.y.is_zero()) return XYZZ<F>::inf();
  F U = dbl(p.y);
  F V = sqr(U);
  F W = mul(U, V);
  F S = mul(p.x, V);
  F X2 = sqr(p.x);
  F M = add(dbl(X2), X2);
  F X3 = sub(sqr(M), dbl(S));
  F Y3 = sub(mul(M, sub(S, X3)), mul(W, p.y));
  return {X3, Y3, V, W};
}

// 2*P (dbl-2008-s-1)
template <class F>
FB_HD XYZZ<F> dbl(const XYZZ<F>& p) {
  if (p.is_inf() || p.y.is_zero()) return XYZZ<F>::inf();
  F U = dbl(p.y);
  F V = sqr(U);
  F W = mul(U, V);
  F S = mul(p.x, V);
  F X2 = sqr(p.x);
  F M = add(dbl(X2), X2);
  F X3 = sub(sqr(M), dbl(S));
  F Y3 = sub(mul(M, sub(S, X3)), mul(W, p.y));
  return {X3, Y3, mul(V, p.zz), mul(W, p.zzz)};
}

// acc + q, q affine (madd-2008-s); handles acc==inf, q==inf, q==+-acc
template <class F>
FB_HD XYZZ<F> add_mixed(const XYZZ<F>& a, const Affine<F>& q) {
  if (q.is_inf()) return a;
  if (a.is_inf()) return {q.x, q.y, F::one(), F::one()};
  F U2 = mul(q.x, a.zz);
  F S2 = mul(q.y, a.zzz);
  F Pp = sub(U2, a.x);
  F R = sub(S2, a.y);
  if (Pp.is_zero()) {
    if (R.is_zero()) return dbl_affine(q);
    return XYZZ<F>::inf();
  }
  F PP = sqr(Pp);
  F PPP = mul(Pp, PP);
  F Q = mul(a.x, PP);
  F X3 = sub(sub(sqr(R), PPP), dbl(Q));
  F Y3 = msub2(R, sub(Q, X3), a.y, PPP);
  return {X3, Y3, mul(a.zz, PP), mul(a.zzz, PPP)};
}

// a + b (add-2008-s)
template <class F>
FB_HD XYZZ<F> add(const XYZZ<F>& a, const XYZZ<F>& b) {
  if (b.is_inf()) return a;
  if (a.is_inf()) return b;
  F U1 = mul(a.x, b.zz);
  F U2 = mul(b.x, a.zz);
  F S1 = mul(a.y, b.zzz);
  F S2 = mul(b.y, a.zzz);
  F Pp = sub(U2, U1);
  F R = sub(S2, S1);
  if (Pp.is_zero()) {
    if (R.is_zero()) return dbl(a);
    return XYZZ<F>::inf();
  }
  F PP = sqr(Pp);
  F PPP = mul(Pp, PP);
  F Q = mul(U1, PP);
  F X3 = sub(sub(sqr(R), PPP), dbl(Q));
  F Y3 = sub(mul(R, sub(Q, X3)), mul(S1, PPP));
  return {X3, Y3, mul(mul(a.zz, b.zz), PP), mul(mul(a.zzz, b.zzz), PPP)};
}

template <class F>
FB_HD Affine<F> neg(const Affine<F>& p) { return {p.x, neg(p.y)}; }
template <class F>
FB_HD XYZZ<F> neg(const XYZZ<F>& p) { return {p.x, neg(p.y), p.zz, p.zzz}; }

template <class F>
FB_HD_COLD Affine<F> to_affine(const XYZZ<F>& p) {
  if (p.is_inf()) return Affine<F>::inf();
  // 1/ZZZ, then 1/ZZ = ZZZ^2 ... avoided: one inversion of ZZ*ZZZ shared
  F t = inv_cold(mul(p.zz, p.zzz));
  F izz = mul(t, p.zzz);
  F izzz = mul(t, p.zz);
  return {mul(p.x, izz), mul(p.y, izzz)};
}

// Out-of-line versions for everything that is not the bucket-accumulation inner loop.
template <class F>
FB_HD_COLD XYZZ<F> add_cold(const XYZZ<F>& a, const XYZZ<F>& b) { return add(a, b); }
template <class F>
FB_HD_COLD XYZZ<F> dbl_cold(const XYZZ<F>& a) { return dbl(a); }
template <class F>
FB_HD_COLD XYZZ<F> add_mixed_cold(const XYZZ<F>& a, const Affine<F>& b) { return add_mixed(a, b); }
template <class F>
FB_HD_COLD F inv_cold(const F& a) { return inv(a); }

// k*P by MSB-first double-and-add; k = 8 canonical (non-Montgomery) 32-bit limbs
template <class F>
FB_HD_COLD XYZZ<F> scalar_mul(const XYZZ<F>& p, const uint32_t* k) {
  XYZZ<F> acc = XYZZ<F>::inf();
  bool started = false;
  for (int i = 255; i >= 0; i--) {
    if (started) acc = dbl_cold(acc);
    if ((k[i >> 5] >> (i & 31)) & 1) {
      acc = add_cold(acc, p);
      started = true;
    }
  }
  return acc;
}

template <class F>
FB_HD bool on_curve(const Affine<F>& p, const F& b) {
  if (p.is_inf()) return true;
  return sqr(p.y) == add(mul(sqr(p.x), p.x), b);
}

using G1Affine = Affine<Fq>;
using G2Affine = Affine<Fq2>;
using G1XYZZ = XYZZ<Fq>;
using G2XYZZ = XYZZ<Fq2>;

FB_HD Fq g1_b() {  // 3 in Montgomery form
  Fq r;
  constexpr uint32_t v[8] = {0x50ad28d7u, 0x7a17caa9u, 0xe15521b9u, 0x1f6ac17au,
                             0x696bd284u, 0x334bea4eu, 0xce179d8eu, 0x2a1f6744u};
  for (int i = 0; i < 8; i++) r.v[i] = v[i];
  return r;
}
FB_HD Fq2 g2_b() {  // 3/(9+u) in Montgomery form
  Fq2 r;
  constexpr uint32_t c0[8] = {0x77b802a8u, 0x3bf938e3u, 0x3633535du, 0x020b1b27u,
                              0x49755260u, 0x26b7edf0u, 0x4384a86du, 0x2514c632u};
  constexpr uint32_t c1[8] = {0xd1dcff67u, 0x38e7ecccu, 0x93ce0d3eu, 0x65f0b37du,
                              0x22ac00aau, 0xd749d0ddu, 0x4a688d4du, 0x0141b9ceu};
  for (int i = 0; i < 8; i++) { r.c0.v[i] = c0[i]; r.c1.v[i] = c1[i]; }
  return r;
}

}  // namespace fb
