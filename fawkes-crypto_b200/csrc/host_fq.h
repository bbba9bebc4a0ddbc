// BN254 Fq / Fq2 on the host with 4 x 64-bit limbs (unsigned __int128): the serial tail of a prove
// (Horner over the MSM bit sums, r- and s-multiples, final affine conversion) runs here while the
// device is still busy.  Same values and memory layout as fb::Fq (8 x u32) / Num<Fq>
// (ff-uint/src/num/mod.rs:21-23); semantics as ff-uint_derive/src/lib.rs:434-490,578-623,836-862.
// The free functions below let the curve templates of ec.cuh instantiate on these types.
#pragma once
#include <cstdint>
#include <cstring>

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
FB_HD constexpr uint64_t modl(int i) { constexpr uint64_t m[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull}; return m[i]; }
constexpr uint64_t INV = 0x87d20782e4866389ull;
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
}  // namespace hfq_detail

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
FB_HD HFq sqr(const HFq& a) { return mul(a, a); }
FB_HD HFq dbl(const HFq& a) { return add(a, a); }
FB_HD HFq neg(const HFq& a) { return sub(HFq::zero(), a); }
FB_HD HFq inv(const HFq& a) {  // a^(p-2)
  using namespace hfq_detail;
  const uint64_t e[4] = {modl(0) - 2, modl(1), modl(2), modl(3)};
  HFq r = HFq::one();
  for (int i = 255; i >= 0; i--) {
    r = sqr(r);
    if ((e[i >> 6] >> (i & 63)) & 1) r = mul(r, a);
  }
  return r;
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
