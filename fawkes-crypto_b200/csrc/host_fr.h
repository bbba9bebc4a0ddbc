// BN254 Fr on the host: 4 x 64-bit Montgomery limbs (unsigned __int128), used by the
// witness/benchmark generator and by setup's scalar stage.  Same semantics as ff.cuh
// (and as ff-uint_derive/src/lib.rs:434-490,578-623,836-862); memory layout identical
// to fb::Fr / Num<Fr>.
#pragma once
#include <cstdint>
#include <cstring>

namespace fb {
namespace hfr {

typedef unsigned __int128 u128;

struct H {
  uint64_t v[4];
};

static const uint64_t MOD[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull,
                                0x30644e72e131a029ull};
static const uint64_t ONE[4] = {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull,
                                0x0e0a77c19a07df2full};
static const uint64_t R2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull,
                               0x0216d0b17f4e44a5ull};
static const uint64_t INV = 0xc2e1f593efffffffull;

inline bool geq_mod(const uint64_t* a) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] > MOD[i]) return true;
    if (a[i] < MOD[i]) return false;
  }
  return true;
}
inline void sub_mod_inplace(uint64_t* a) {
  u128 bw = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a[i] - MOD[i] - (uint64_t)bw;
    a[i] = (uint64_t)d;
    bw = (d >> 64) & 1;
  }
}
inline H zero() { H r; memset(r.v, 0, 32); return r; }
inline H one() { H r; memcpy(r.v, ONE, 32); return r; }
inline bool is_zero(const H& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
inline bool eq(const H& a, const H& b) { return !memcmp(a.v, b.v, 32); }

inline H add(const H& a, const H& b) {
  H r;
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a.v[i] + b.v[i];
    r.v[i] = (uint64_t)c;
    c >>= 64;
  }
  if (geq_mod(r.v)) sub_mod_inplace(r.v);
  return r;
}
inline H sub(const H& a, const H& b) {
  H r;
  u128 bw = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a.v[i] - b.v[i] - (uint64_t)bw;
    r.v[i] = (uint64_t)d;
    bw = (d >> 64) & 1;
  }
  if (bw) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)r.v[i] + MOD[i];
      r.v[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  return r;
}
inline H mul(const H& a, const H& b) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a.v[j] * b.v[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (uint64_t)c;
    t[5] = (uint64_t)(c >> 64);
    uint64_t q = t[0] * INV;
    c = ((u128)q * MOD[0] + t[0]) >> 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)q * MOD[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = t[5] + (uint64_t)(c >> 64);
  }
  H r;
  memcpy(r.v, t, 32);
  if (geq_mod(r.v)) sub_mod_inplace(r.v);
  return r;
}
inline H to_mont(const H& a) { H r2; memcpy(r2.v, R2, 32); return mul(a, r2); }
inline H from_mont(const H& a) { H o = zero(); o.v[0] = 1; return mul(a, o); }
inline H pow(const H& a, const uint64_t* e, int limbs) {
  H r = one();
  for (int i = limbs * 64 - 1; i >= 0; i--) {
    r = mul(r, r);
    if ((e[i >> 6] >> (i & 63)) & 1) r = mul(r, a);
  }
  return r;
}
inline H inv(const H& a) {
  uint64_t e[4] = {MOD[0] - 2, MOD[1], MOD[2], MOD[3]};
  return pow(a, e, 4);
}

}  // namespace hfr
}  // namespace fb
