// Shared pieces of the C++ CPU restatement: BN254 fields (4 x 64-bit limbs, unsigned __int128),
// Jacobian points, the thread helpers and bellman's point byte format.
//
// TEST INFRASTRUCTURE ONLY (see cpu_prover.cpp for the import rule and the parity status).
//   field mul/add/sub        ff-uint_derive/src/lib.rs:434-490,578-623,836-862 (same results)
//   moduli                   fawkes-crypto/src/engines/bn256/mod.rs:13,23
//   point byte formats       SURVEY.md App. B
#pragma once
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

typedef unsigned __int128 u128;
typedef uint64_t u64;

// ------------------------------------------------------------------ fields ---
struct FrP {
  static constexpr u64 MOD[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  static constexpr u64 R[4] = {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full};
  static constexpr u64 R2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
  static constexpr u64 INV = 0xc2e1f593efffffffull;
};
struct FqP {
  static constexpr u64 MOD[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  static constexpr u64 R[4] = {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full};
  static constexpr u64 R2[4] = {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full};
  static constexpr u64 INV = 0x87d20782e4866389ull;
};
constexpr u64 FrP::MOD[4]; constexpr u64 FrP::R[4]; constexpr u64 FrP::R2[4];
constexpr u64 FqP::MOD[4]; constexpr u64 FqP::R[4]; constexpr u64 FqP::R2[4];

template <class P>
struct Fp {
  u64 v[4];
  static Fp zero() { Fp r; memset(r.v, 0, 32); return r; }
  static Fp one() { Fp r; memcpy(r.v, P::R, 32); return r; }
  bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
  bool operator==(const Fp& o) const { return !memcmp(v, o.v, 32); }
  static bool geq(const u64* a) {
    for (int i = 3; i >= 0; i--) {
      if (a[i] > P::MOD[i]) return true;
      if (a[i] < P::MOD[i]) return false;
    }
    return true;
  }
  static void subm(u64* a) {
    u64 bw = 0;
    for (int i = 0; i < 4; i++) {
      u128 d = (u128)a[i] - P::MOD[i] - bw;
      a[i] = (u64)d;
      bw = (u64)(d >> 64) & 1;
    }
  }
  Fp operator+(const Fp& b) const {
    Fp r;
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)v[i] + b.v[i]; r.v[i] = (u64)c; c >>= 64; }
    if (geq(r.v)) subm(r.v);
    return r;
  }
  Fp operator-(const Fp& b) const {
    Fp r;
    u64 bw = 0;
    for (int i = 0; i < 4; i++) {
      u128 d = (u128)v[i] - b.v[i] - bw;
      r.v[i] = (u64)d;
      bw = (u64)(d >> 64) & 1;
    }
    if (bw) {
      u128 c = 0;
      for (int i = 0; i < 4; i++) { c += (u128)r.v[i] + P::MOD[i]; r.v[i] = (u64)c; c >>= 64; }
    }
    return r;
  }
  Fp neg() const { return is_zero() ? *this : zero() - *this; }
  Fp dbl() const { return *this + *this; }
  // schoolbook 4x4 product then 4-round Montgomery reduction (ff-uint_derive's SOS shape)
  Fp operator*(const Fp& b) const {
    u64 t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
      u64 carry = 0;
      for (int j = 0; j < 4; j++) {
        u128 x = (u128)v[i] * b.v[j] + t[i + j] + carry;
        t[i + j] = (u64)x;
        carry = (u64)(x >> 64);
      }
      t[i + 4] = carry;
    }
    u64 carry2 = 0;
    for (int i = 0; i < 4; i++) {
      u64 k = t[i] * P::INV;
      u128 x = (u128)k * P::MOD[0] + t[i];
      u64 carry = (u64)(x >> 64);
      for (int j = 1; j < 4; j++) {
        x = (u128)k * P::MOD[j] + t[i + j] + carry;
        t[i + j] = (u64)x;
        carry = (u64)(x >> 64);
      }
      x = (u128)t[i + 4] + carry2 + carry;
      t[i + 4] = (u64)x;
      carry2 = (u64)(x >> 64);
    }
    Fp r;
    memcpy(r.v, t + 4, 32);
    if (geq(r.v)) subm(r.v);
    return r;
  }
  Fp sqr() const { return *this * *this; }
  Fp to_mont() const { Fp r2; memcpy(r2.v, P::R2, 32); return *this * r2; }
  Fp from_mont() const { Fp o = zero(); o.v[0] = 1; return *this * o; }
  Fp pow(const u64* e, int limbs) const {
    Fp r = one();
    for (int i = limbs * 64 - 1; i >= 0; i--) {
      r = r.sqr();
      if ((e[i >> 6] >> (i & 63)) & 1) r = r * *this;
    }
    return r;
  }
  Fp inv() const {
    u64 e[4] = {P::MOD[0] - 2, P::MOD[1], P::MOD[2], P::MOD[3]};
    return pow(e, 4);
  }
};
typedef Fp<FrP> Fr;
typedef Fp<FqP> Fq;

struct Fq2 {
  Fq c0, c1;
  static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
  static Fq2 one() { return {Fq::one(), Fq::zero()}; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
  Fq2 operator+(const Fq2& b) const { return {c0 + b.c0, c1 + b.c1}; }
  Fq2 operator-(const Fq2& b) const { return {c0 - b.c0, c1 - b.c1}; }
  Fq2 neg() const { return {c0.neg(), c1.neg()}; }
  Fq2 dbl() const { return {c0.dbl(), c1.dbl()}; }
  Fq2 operator*(const Fq2& b) const {
    Fq aa = c0 * b.c0, bb = c1 * b.c1;
    Fq o = (c0 + c1) * (b.c0 + b.c1);
    return {aa - bb, o - aa - bb};
  }
  Fq2 sqr() const {
    Fq ab = c0 * c1;
    return {(c0 + c1) * (c0 - c1), ab.dbl()};
  }
  Fq2 inv() const {
    Fq t = (c0.sqr() + c1.sqr()).inv();
    return {c0 * t, (c1 * t).neg()};
  }
};

// ------------------------------------------------------- Jacobian points ---
template <class F>
struct Aff {
  F x, y;
  bool inf;
};
template <class F>
struct Jac {
  F x, y, z;
  static Jac zero() { return {F::zero(), F::one(), F::zero()}; }
  bool is_zero() const { return z.is_zero(); }
  void dbl() {  // dbl-2009-l
    if (is_zero()) return;
    F a = x.sqr(), b = y.sqr(), c = b.sqr();
    F d = ((x + b).sqr() - a - c).dbl();
    F e = a.dbl() + a, f = e.sqr();
    z = (z * y).dbl();
    x = f - d.dbl();
    y = e * (d - x) - c.dbl().dbl().dbl();
  }
  void add(const Jac& o) {  // add-2007-bl
    if (is_zero()) { *this = o; return; }
    if (o.is_zero()) return;
    F z1z1 = z.sqr(), z2z2 = o.z.sqr();
    F u1 = x * z2z2, u2 = o.x * z1z1;
    F s1 = y * o.z * z2z2, s2 = o.y * z * z1z1;
    if (u1 == u2 && s1 == s2) { dbl(); return; }
    F h = u2 - u1, i = h.dbl().sqr(), j = h * i;
    F r = (s2 - s1).dbl(), v = u1 * i;
    F nx = r.sqr() - j - v.dbl();
    F ny = r * (v - nx) - (s1 * j).dbl();
    z = ((z + o.z).sqr() - z1z1 - z2z2) * h;
    x = nx;
    y = ny;
  }
  void add_mixed(const Aff<F>& o) {  // madd-2007-bl
    if (o.inf) return;
    if (is_zero()) { x = o.x; y = o.y; z = F::one(); return; }
    F z1z1 = z.sqr();
    F u2 = o.x * z1z1, s2 = o.y * z * z1z1;
    if (x == u2 && y == s2) { dbl(); return; }
    F h = u2 - x, hh = h.sqr(), i = hh.dbl().dbl(), j = h * i;
    F r = (s2 - y).dbl(), v = x * i;
    F nx = r.sqr() - j - v.dbl();
    F ny = r * (v - nx) - (y * j).dbl();
    z = (z + h).sqr() - z1z1 - hh;
    x = nx;
    y = ny;
  }
  void mul_assign(const u64* k) {  // canonical scalar, MSB first
    Jac res = zero();
    bool found = false;
    for (int i = 255; i >= 0; i--) {
      if (found) res.dbl();
      if ((k[i >> 6] >> (i & 63)) & 1) { found = true; res.add(*this); }
    }
    *this = res;
  }
  Aff<F> to_affine() const {
    if (is_zero()) return {F::zero(), F::zero(), true};
    F zi = z.inv(), zi2 = zi.sqr();
    return {x * zi2, y * zi2 * zi, false};
  }
};

// ------------------------------------------------------------- threading ---
static void parallel_for(int nthreads, size_t n, const std::function<void(size_t, size_t)>& f) {
  if (nthreads <= 1 || n < 2) { f(0, n); return; }
  std::vector<std::thread> th;
  size_t chunk = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; t++) {
    size_t lo = t * chunk, hi = std::min(n, lo + chunk);
    if (lo >= hi) break;
    th.emplace_back([=, &f] { f(lo, hi); });
  }
  for (auto& t : th) t.join();
}
static void run_tasks(int nthreads, std::vector<std::function<void()>>& tasks) {
  std::atomic<size_t> next(0);
  auto worker = [&] {
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= tasks.size()) return;
      tasks[i]();
    }
  };
  if (nthreads <= 1) { worker(); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) th.emplace_back(worker);
  for (auto& t : th) t.join();
}

// ------------------------------------------------------------ byte formats ---
static Fq fq_from_be(const uint8_t* b) {
  Fq r;
  for (int i = 0; i < 4; i++) {
    u64 v = 0;
    for (int j = 0; j < 8; j++) v = (v << 8) | b[(3 - i) * 8 + j];
    r.v[i] = v;
  }
  return r.to_mont();
}
static Aff<Fq> g1_from_be(const uint8_t* b) {
  if (b[0] & 0x40) return {Fq::zero(), Fq::zero(), true};
  return {fq_from_be(b), fq_from_be(b + 32), false};
}
static Aff<Fq2> g2_from_be(const uint8_t* b) {
  if (b[0] & 0x40) return {Fq2::zero(), Fq2::zero(), true};
  Fq x1 = fq_from_be(b), x0 = fq_from_be(b + 32), y1 = fq_from_be(b + 64), y0 = fq_from_be(b + 96);
  return {{x0, x1}, {y0, y1}, false};
}
static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }


// ---------------------------------------------------------------- domain ---
static const u64 ROOT_OF_UNITY[4] = {0x9632c7c5b639feb8ull, 0x985ce3400d0ff299ull, 0xb2dd880001b0ecd8ull, 0x1d69070d6d98ce29ull};
static const u64 GEN7[4] = {0x3057819e4fffffdbull, 0x307f6d866832bb01ull, 0x5c65ec9f484e3a89ull, 0x0180a96573d3d9f8ull};

static Fr fr_pow_u64(Fr a, u64 e) { return a.pow(&e, 1); }

static void fft(std::vector<Fr>& a, const Fr& omega, int exp, int nthreads) {
  const size_t n = a.size();
  for (size_t k = 0; k < n; k++) {
    size_t rk = 0;
    for (int b = 0; b < exp; b++) rk |= ((k >> b) & 1) << (exp - 1 - b);
    if (k < rk) std::swap(a[k], a[rk]);
  }
  size_t m = 1;
  for (int s = 0; s < exp; s++) {
    Fr w_m = fr_pow_u64(omega, n / (2 * m));
    const size_t groups = n / (2 * m);
    if (groups >= (size_t)nthreads * 4 || nthreads <= 1) {
      parallel_for(nthreads, groups, [&](size_t lo, size_t hi) {
        for (size_t g = lo; g < hi; g++) {
          size_t k = g * 2 * m;
          Fr w = Fr::one();
          for (size_t j = 0; j < m; j++) {
            Fr t = a[k + j + m] * w;
            a[k + j + m] = a[k + j] - t;
            a[k + j] = a[k + j] + t;
            w = w * w_m;
          }
        }
      });
    } else {  // few large groups: split the j range
      for (size_t g = 0; g < groups; g++) {
        size_t k = g * 2 * m;
        parallel_for(nthreads, m, [&](size_t lo, size_t hi) {
          Fr w = fr_pow_u64(w_m, lo);
          for (size_t j = lo; j < hi; j++) {
            Fr t = a[k + j + m] * w;
            a[k + j + m] = a[k + j] - t;
            a[k + j] = a[k + j] + t;
            w = w * w_m;
          }
        });
      }
    }
    m *= 2;
  }
}
static void distribute_powers(std::vector<Fr>& a, const Fr& g, int nthreads) {
  parallel_for(nthreads, a.size(), [&](size_t lo, size_t hi) {
    Fr u = fr_pow_u64(g, lo);
    for (size_t i = lo; i < hi; i++) { a[i] = a[i] * u; u = u * g; }
  });
}
static void scale(std::vector<Fr>& a, const Fr& s, int nthreads) {
  parallel_for(nthreads, a.size(), [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; i++) a[i] = a[i] * s; });
}

