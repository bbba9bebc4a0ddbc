// BN254 Fq / Fq2 on the host with 4 x 64-bit limbs (unsigned __int128): the serial tail of a prove
// (Horner over the MSM bit sums, r- and s-multiples, final affine conversion) runs here while the
// device is still busy.  Same values and memory layout as fb::Fq (8 x u32) / Num<Fq>
// (ff-uint/src/num/mod.rs:21-23); semantics as ff-uint_derive/src/lib.rs:434-490,578-623,836-862.
// The free functions below let the curve templates of ec.cuh instantiate on these types.
#pragma once
#include <cstdint>
#include <cstring>

#if !defined(__CUDA_ARCH__) && defined(__x86_64__)
#include <immintrin.h>
#define FB_HFQ_X86 1
#else
#define FB_HFQ_X86 0
#endif

#include "ff.cuh"

namespace fb {

struct HFq {
  uint64_t v[4];
  FB_HD static HFq zero() { HFq r; memset(r.v, 0, 32); return r; }
  FB_HD static HFq one() {
    HFq r;
    const uint64_t o[4] = {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full};
    memcpy(r.v, o, 32);
    return r;
  }
  FB_HD bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
  FB_HD bool operator==(const HFq& o) const { return ((v[0] ^ o.v[0]) | (v[1] ^ o.v[1]) | (v[2] ^ o.v[2]) | (v[3] ^ o.v[3])) == 0; }
  FB_HD bool operator!=(const HFq& o) const { return !(*this == o); }
  FB_HD static HFq from(const Fq& x) { HFq r; memcpy(r.v, x.v, 32); return r; }
  FB_HD Fq to() const { Fq r; memcpy(r.v, v, 32); return r; }
};

namespace hfq_detail {
typedef unsigned __int128 u128;
typedef unsigned long long ull;
FB_HD constexpr uint64_t modl(int i) { constexpr uint64_t m[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull}; return m[i]; }
constexpr uint64_t INV = 0x87d20782e4866389ull;
#if FB_HFQ_X86
// r in [0, 2p) -> [0, p): subtract p, keep the difference unless it borrowed (cmov, no branch)
inline void reduce_once(ull& r0, ull& r1, ull& r2, ull& r3) {
  ull t0, t1, t2, t3;
  unsigned char bw = _subborrow_u64(0, r0, modl(0), &t0);
  bw = _subborrow_u64(bw, r1, modl(1), &t1);
  bw = _subborrow_u64(bw, r2, modl(2), &t2);
  bw = _subborrow_u64(bw, r3, modl(3), &t3);
  r0 = bw ? r0 : t0;
  r1 = bw ? r1 : t1;
  r2 = bw ? r2 : t2;
  r3 = bw ? r3 : t3;
}
#else
FB_HD bool geq(const uint64_t* a) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] > modl(i)) return true;
    if (a[i] < modl(i)) return false;
  }
  return true;
}
FB_HD void subm(uint64_t* a) {
  uint64_t bw = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a[i] - modl(i) - bw;
    a[i] = (uint64_t)d;
    bw = (uint64_t)(d >> 64) & 1;
  }
}
#endif
}  // namespace hfq_detail

#if FB_HFQ_X86
// x86-64 host build: add/adc and sub/sbb chains through the carry intrinsics, a fully unrolled Montgomery product
// (CIOS without the extra carry word: p < 2^254).  The tower arithmetic of the pairing (verify.cu) is dominated by
// additions; the portable loops below compile to 4x slower code for them.
FB_HD HFq add(const HFq& a, const HFq& b) {
  using namespace hfq_detail;
  ull r0, r1, r2, r3;
  unsigned char c = _addcarry_u64(0, a.v[0], b.v[0], &r0);
  c = _addcarry_u64(c, a.v[1], b.v[1], &r1);
  c = _addcarry_u64(c, a.v[2], b.v[2], &r2);
  c = _addcarry_u64(c, a.v[3], b.v[3], &r3);  // no carry out: 2p < 2^255
  reduce_once(r0, r1, r2, r3);
  HFq r;
  r.v[0] = r0; r.v[1] = r1; r.v[2] = r2; r.v[3] = r3;
  return r;
}
FB_HD HFq sub(const HFq& a, const HFq& b) {
  using namespace hfq_detail;
  ull r0, r1, r2, r3;
  unsigned char bw = _subborrow_u64(0, a.v[0], b.v[0], &r0);
  bw = _subborrow_u64(bw, a.v[1], b.v[1], &r1);
  bw = _subborrow_u64(bw, a.v[2], b.v[2], &r2);
  bw = _subborrow_u64(bw, a.v[3], b.v[3], &r3);
  const ull m = 0 - (ull)bw;  // add p back when the difference went negative
  unsigned char c = _addcarry_u64(0, r0, modl(0) & m, &r0);
  c = _addcarry_u64(c, r1, modl(1) & m, &r1);
  c = _addcarry_u64(c, r2, modl(2) & m, &r2);
  c = _addcarry_u64(c, r3, modl(3) & m, &r3);
  HFq r;
  r.v[0] = r0; r.v[1] = r1; r.v[2] = r2; r.v[3] = r3;
  return r;
}
FB_HD HFq mul(const HFq& a, const HFq& b) {
  using namespace hfq_detail;
  ull t0 = 0, t1 = 0, t2 = 0, t3 = 0;
#define FB_HFQ_ROUND(bi)                                                         \
  {                                                                              \
    u128 x = (u128)a.v[0] * (bi) + t0;                                           \
    ull A = (ull)(x >> 64);                                                      \
    const ull m = (ull)x * INV;                                                  \
    u128 y = (u128)m * modl(0) + (ull)x;                                         \
    ull C = (ull)(y >> 64);                                                      \
    x = (u128)a.v[1] * (bi) + t1 + A; A = (ull)(x >> 64);                        \
    y = (u128)m * modl(1) + (ull)x + C; C = (ull)(y >> 64); t0 = (ull)y;         \
    x = (u128)a.v[2] * (bi) + t2 + A; A = (ull)(x >> 64);                        \
    y = (u128)m * modl(2) + (ull)x + C; C = (ull)(y >> 64); t1 = (ull)y;         \
    x = (u128)a.v[3] * (bi) + t3 + A; A = (ull)(x >> 64);                        \
    y = (u128)m * modl(3) + (ull)x + C; C = (ull)(y >> 64); t2 = (ull)y;         \
    t3 = C + A;                                                                  \
  }
  FB_HFQ_ROUND(b.v[0]) FB_HFQ_ROUND(b.v[1]) FB_HFQ_ROUND(b.v[2]) FB_HFQ_ROUND(b.v[3])
#undef FB_HFQ_ROUND
  reduce_once(t0, t1, t2, t3);  // the Montgomery product of two reduced operands is < 2p
  HFq r;
  r.v[0] = t0; r.v[1] = t1; r.v[2] = t2; r.v[3] = t3;
  return r;
}
#else
FB_HD HFq add(const HFq& a, const HFq& b) {
  using namespace hfq_detail;
  HFq r;
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a.v[i] + b.v[i]; r.v[i] = (uint64_t)c; c >>= 64; }
  if (geq(r.v)) subm(r.v);
  return r;
}
FB_HD HFq sub(const HFq& a, const HFq& b) {
  using namespace hfq_detail;
  HFq r;
  uint64_t bw = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a.v[i] - b.v[i] - bw;
    r.v[i] = (uint64_t)d;
    bw = (uint64_t)(d >> 64) & 1;
  }
  if (bw) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)r.v[i] + modl(i); r.v[i] = (uint64_t)c; c >>= 64; }
  }
  return r;
}
FB_HD HFq mul(const HFq& a, const HFq& b) {
  using namespace hfq_detail;
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a.v[j] * b.v[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[4];
    t[4] = (uint64_t)c;
    t[5] = (uint64_t)(c >> 64);
    uint64_t q = t[0] * INV;
    c = ((u128)q * modl(0) + t[0]) >> 64;
    for (int j = 1; j < 4; j++) { c += (u128)q * modl(j) + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = t[5] + (uint64_t)(c >> 64);
  }
  HFq r;
  memcpy(r.v, t, 32);
  if (geq(r.v)) subm(r.v);
  return r;
}
#endif
FB_HD HFq sqr(const HFq& a) { return mul(a, a); }
FB_HD HFq dbl(const HFq& a) { return add(a, a); }
FB_HD HFq neg(const HFq& a) { return sub(HFq::zero(), a); }
FB_HD HFq inv_fermat(const HFq& a) {  // a^(p-2); the yardstick of inv() in fb_test_pairing
  using namespace hfq_detail;
  const uint64_t e[4] = {modl(0) - 2, modl(1), modl(2), modl(3)};
  HFq r = HFq::one();
  for (int i = 255; i >= 0; i--) {
    r = sqr(r);
    if ((e[i >> 6] >> (i & 63)) & 1) r = mul(r, a);
  }
  return r;
}
// Binary extended Euclid on the Montgomery residue (Hankerson et al., Alg. 2.22): 2.4x faster than a^(p-2) on
// the host (no multiplications, ~380 shift / subtract steps).  The input aR gives (aR)^-1; one Montgomery product
// with R^3 turns that into a^-1 R.  inv(0) = 0, as with Fermat.
FB_HD HFq inv(const HFq& a) {
  using namespace hfq_detail;
  if (a.is_zero()) return a;
  uint64_t u[4], v[4], x1[4] = {1, 0, 0, 0}, x2[4] = {0, 0, 0, 0};
  for (int i = 0; i < 4; i++) { u[i] = a.v[i]; v[i] = modl(i); }
  auto is_one = [](const uint64_t* x) { return x[0] == 1 && (x[1] | x[2] | x[3]) == 0; };
  auto shr1 = [](uint64_t* x) {
    x[0] = (x[0] >> 1) | (x[1] << 63);
    x[1] = (x[1] >> 1) | (x[2] << 63);
    x[2] = (x[2] >> 1) | (x[3] << 63);
    x[3] >>= 1;
  };
  auto half_mod = [&](uint64_t* x) {  // x / 2 mod p for x < p (x + p < 2^255)
    if (x[0] & 1) {
      u128 c = 0;
      for (int i = 0; i < 4; i++) { c += (u128)x[i] + modl(i); x[i] = (uint64_t)c; c >>= 64; }
    }
    shr1(x);
  };
  auto geq_ = [](const uint64_t* x, const uint64_t* y) {
    for (int i = 3; i >= 0; i--)
      if (x[i] != y[i]) return x[i] > y[i];
    return true;
  };
  auto sub_ = [](uint64_t* x, const uint64_t* y) {  // x -= y, returns the borrow
    uint64_t bw = 0;
    for (int i = 0; i < 4; i++) {
      u128 d = (u128)x[i] - y[i] - bw;
      x[i] = (uint64_t)d;
      bw = (uint64_t)(d >> 64) & 1;
    }
    return bw;
  };
  auto sub_mod = [&](uint64_t* x, const uint64_t* y) {
    if (sub_(x, y)) {
      u128 c = 0;
      for (int i = 0; i < 4; i++) { c += (u128)x[i] + modl(i); x[i] = (uint64_t)c; c >>= 64; }
    }
  };
  // the loop needs 0 < u < p (gcd(u, p) = 1); every HFq in the library is reduced, but a limb pattern >= p must
  // not hang the caller: bring it below p first (2^256 < 6p)
  for (int k = 0; k < 6; k++)
    if (geq_(u, v)) sub_(u, v);
  if ((u[0] | u[1] | u[2] | u[3]) == 0) return HFq::zero();
  while (!is_one(u) && !is_one(v)) {
    while (!(u[0] & 1)) { shr1(u); half_mod(x1); }
    while (!(v[0] & 1)) { shr1(v); half_mod(x2); }
    if (geq_(u, v)) { sub_(u, v); sub_mod(x1, x2); }
    else { sub_(v, u); sub_mod(x2, x1); }
  }
  HFq y, r3;
  const uint64_t R3[4] = {0xb1cd6dafda1530dfull, 0x62f210e6a7283db6ull, 0xef7f0b0c0ada0afbull, 0x20fd6e902d592544ull};  // R^3 mod p
  for (int i = 0; i < 4; i++) { y.v[i] = is_one(u) ? x1[i] : x2[i]; r3.v[i] = R3[i]; }
  return mul(y, r3);
}

struct HFq2 {
  HFq c0, c1;
  FB_HD static HFq2 zero() { return {HFq::zero(), HFq::zero()}; }
  FB_HD static HFq2 one() { return {HFq::one(), HFq::zero()}; }
  FB_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  FB_HD bool operator==(const HFq2& o) const { return c0 == o.c0 && c1 == o.c1; }
  FB_HD bool operator!=(const HFq2& o) const { return !(*this == o); }
  FB_HD static HFq2 from(const Fq2& x) { return {HFq::from(x.c0), HFq::from(x.c1)}; }
  FB_HD Fq2 to() const { return {c0.to(), c1.to()}; }
};
FB_HD HFq2 add(const HFq2& a, const HFq2& b) { return {add(a.c0, b.c0), add(a.c1, b.c1)}; }
FB_HD HFq2 sub(const HFq2& a, const HFq2& b) { return {sub(a.c0, b.c0), sub(a.c1, b.c1)}; }
FB_HD HFq2 dbl(const HFq2& a) { return {dbl(a.c0), dbl(a.c1)}; }
FB_HD HFq2 neg(const HFq2& a) { return {neg(a.c0), neg(a.c1)}; }
FB_HD HFq2 mul(const HFq2& a, const HFq2& b) {
  HFq t0 = mul(a.c0, b.c0), t1 = mul(a.c1, b.c1);
  HFq t2 = mul(add(a.c0, a.c1), add(b.c0, b.c1));
  return {sub(t0, t1), sub(sub(t2, t0), t1)};
}
FB_HD HFq2 sqr(const HFq2& a) {
  HFq t = mul(a.c0, a.c1);
  return {mul(add(a.c0, a.c1), sub(a.c0, a.c1)), dbl(t)};
}
FB_HD HFq2 inv(const HFq2& a) {
  HFq n = inv(add(sqr(a.c0), sqr(a.c1)));
  return {mul(a.c0, n), neg(mul(a.c1, n))};
}

}  // namespace fb
