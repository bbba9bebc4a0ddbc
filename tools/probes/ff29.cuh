// Lazy 9 x 29-bit limb arithmetic for BN254 Fq / Fr: the "un-chained" multiplier.
//
// Why: on sm_100a a carry-chained wide MAC (IMAD.WIDE.U32.X, what mad.lo.cc/madc.hi.cc compile to)
// issues at HALF the rate of a plain IMAD.WIDE.U32 (8.5e12 vs 1.72e13 MAC32/s measured,
// profiles/r01_rate_probe.txt).  With 29-bit limbs a column of 9 + 9 products (< 2^62.2) fits a 64-bit
// accumulator, so a Montgomery product is 162 plain `acc += a*b` MACs and a handful of shifts/masks,
// no carry flags at all.
//
// Representation ("internal form"): value x is held as x * 2^261 mod p (Montgomery radix 2^(9*29)),
// limbs < 2^29, value in [0, 2p) ("weakly reduced").  Because 2^261 > 128 p, a product of two such
// values is again < 2p without any conditional subtraction.  Relation to the reference's memory form
// (x * 2^256, ff-uint/src/num/mod.rs:21-23):  internal * internal -> internal,
// internal * standard -> standard, so kernels mix the two without conversions where they can.
// Results are unique field elements, hence bit-identical to the reference after conversion.
#pragma once
#include "ff.cuh"
#include "ff29_consts.h"

namespace fb {

struct Fq29Cfg {
  using Base = FqCfg;
  static constexpr uint32_t PINV = FB_C29_FQ_PINV;
  FB_HD static constexpr uint32_t p(int i) { constexpr uint32_t v[9] = FB_C29_FQ_P; return v[i]; }
  FB_HD static constexpr uint32_t p2(int i) { constexpr uint32_t v[9] = FB_C29_FQ_2P; return v[i]; }
  FB_HD static constexpr uint32_t one(int i) { constexpr uint32_t v[9] = FB_C29_FQ_ONE; return v[i]; }
  FB_HD static constexpr uint32_t to_std(int i) { constexpr uint32_t v[9] = FB_C29_FQ_TO_STD; return v[i]; }
  FB_HD static constexpr uint32_t to_int_w(int i) { constexpr uint32_t v[8] = FB_C29_FQ_TO_INT_W; return v[i]; }
};
struct Fr29Cfg {
  using Base = FrCfg;
  static constexpr uint32_t PINV = FB_C29_FR_PINV;
  FB_HD static constexpr uint32_t p(int i) { constexpr uint32_t v[9] = FB_C29_FR_P; return v[i]; }
  FB_HD static constexpr uint32_t p2(int i) { constexpr uint32_t v[9] = FB_C29_FR_2P; return v[i]; }
  FB_HD static constexpr uint32_t one(int i) { constexpr uint32_t v[9] = FB_C29_FR_ONE; return v[i]; }
  FB_HD static constexpr uint32_t to_std(int i) { constexpr uint32_t v[9] = FB_C29_FR_TO_STD; return v[i]; }
  FB_HD static constexpr uint32_t to_int_w(int i) { constexpr uint32_t v[8] = FB_C29_FR_TO_INT_W; return v[i]; }
};

constexpr uint32_t M29 = (1u << 29) - 1;


// acc += x * y as ONE plain wide MAC (no carry flags).  The asm form keeps ptxas from turning a
// multiply by a modulus limb into an immediate form that needs an extra add per MAC.
FB_HD void mac29(uint64_t& acc, uint32_t x, uint32_t y) { acc += (uint64_t)x * y; }
// A modulus limb the optimiser cannot see through: otherwise nvcc rewrites u64(q) * constant as a
// 64-bit constant multiply (an extra add per MAC); as a plain register it stays one IMAD.WIDE.
#if defined(__CUDACC__)
static __device__ uint32_t g_fb_zero29;  // always 0; read at run time so modulus limbs live in registers
#endif
FB_HD uint32_t reg29(uint32_t v) {
#if defined(__CUDA_ARCH__) && defined(FB_P_IN_REGS)
  return v | g_fb_zero29;
#else
  return v;
#endif
}
FB_HD uint32_t opaque29(uint32_t v) {
#if defined(__CUDA_ARCH__)
  asm("" : "+r"(v));
#endif
  return v;
}

template <class C>
struct Fl {  // limbs < 2^29, value < 2p
  uint32_t l[9];
  using Cfg = C;
  FB_HD static Fl zero() {
    Fl r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.l[i] = 0;
    return r;
  }
  FB_HD static Fl one() {
    Fl r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.l[i] = C::one(i);
    return r;
  }
  // value == 0 mod p  <=>  value in {0, p}
  FB_HD bool is_zero() const {
    uint32_t z = 0, e = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) { z |= l[i]; e |= l[i] ^ C::p(i); }
    return z == 0 || e == 0;
  }
};

// r = (s >= 2p) ? s - 2p : s   for normalized s < 4p
template <class C>
FB_HD Fl<C> cond_sub_2p(const Fl<C>& s) {
  Fl<C> d;
  int32_t c = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    int32_t t = (int32_t)s.l[i] - (int32_t)C::p2(i) + c;
    d.l[i] = (uint32_t)t & M29;
    c = t >> 29;
  }
  // the top limb carries no mask semantics: value negative <=> final borrow
  Fl<C> r;
#pragma unroll
  for (int i = 0; i < 9; i++) r.l[i] = c < 0 ? s.l[i] : d.l[i];
  return r;
}

template <class C>
FB_HD Fl<C> add(const Fl<C>& a, const Fl<C>& b) {
  Fl<C> s;
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    uint32_t t = a.l[i] + b.l[i] + c;
    s.l[i] = i < 8 ? (t & M29) : t;
    c = t >> 29;
  }
  return cond_sub_2p(s);
}

template <class C>
FB_HD Fl<C> sub(const Fl<C>& a, const Fl<C>& b) {  // a - b + 2p, then weak reduce
  Fl<C> s;
  int32_t c = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    int32_t t = (int32_t)a.l[i] - (int32_t)b.l[i] + (int32_t)C::p2(i) + c;
    s.l[i] = i < 8 ? ((uint32_t)t & M29) : (uint32_t)t;
    c = t >> 29;
  }
  return cond_sub_2p(s);
}

template <class C>
FB_HD Fl<C> dbl(const Fl<C>& a) { return add(a, a); }
template <class C>
FB_HD Fl<C> neg(const Fl<C>& a) { return sub(Fl<C>::zero(), a); }

// Montgomery product a*b*2^-261, operand scanning ("row-wise") on 64-bit column accumulators
// without carry flags.  Row i issues nine MACs that share the multiplier b_i and hit nine different
// accumulators, then nine that share q_i: every MAC is a plain IMAD.WIDE.U32 whose multiplier
// operand is reused from the operand cache and whose accumulators are independent (ILP 9), the
// shape that reaches the 64 MAC/clk/SM rate of the integer pipe (profiles/r01_rate_probe.txt).
// Inputs: limbs < 2^30 (one un-normalised sum may be fed in), value(a)*value(b) < 128 p^2.
// Column bound: 9 a*b + 9 q*p + carry < 2^63.6.
template <class C>
FB_HD Fl<C> mul(const Fl<C>& a, const Fl<C>& b) {
  uint64_t c[19];
  uint32_t pl[9];
  Fl<C> r;
#pragma unroll
  for (int i = 0; i < 19; i++) c[i] = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) pl[i] = reg29(C::p(i));
#pragma unroll
  for (int i = 0; i < 9; i++) {
#pragma unroll
    for (int j = 0; j < 9; j++) mac29(c[i + j], a.l[j], b.l[i]);
    // opaque: keeps q a 32-bit register (else nvcc widens q*p into a 64x64 multiply)
    const uint32_t q = opaque29(((uint32_t)c[i] * C::PINV) & M29);
#pragma unroll
    for (int j = 0; j < 9; j++) mac29(c[i + j], q, pl[j]);
    c[i + 1] += c[i] >> 29;  // low 29 bits of c[i] are zero now
  }
#pragma unroll
  for (int k = 9; k < 18; k++) {
    r.l[k - 9] = opaque29((uint32_t)c[k] & M29);
    c[k + 1] += c[k] >> 29;
  }
  return r;
}

// a^2: row i multiplies a_i with itself and with the doubled a_j, j > i (45 products instead of 81)
template <class C>
FB_HD Fl<C> sqr(const Fl<C>& a) {
  uint64_t c[19];
  uint32_t pl[9], a2[9];
  Fl<C> r;
#pragma unroll
  for (int i = 0; i < 19; i++) c[i] = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) { pl[i] = reg29(C::p(i)); a2[i] = a.l[i] << 1; }
#pragma unroll
  for (int i = 0; i < 9; i++) {
    mac29(c[2 * i], a.l[i], a.l[i]);
#pragma unroll
    for (int j = i + 1; j < 9; j++) mac29(c[i + j], a2[j], a.l[i]);
    const uint32_t q = opaque29(((uint32_t)c[i] * C::PINV) & M29);
#pragma unroll
    for (int j = 0; j < 9; j++) mac29(c[i + j], q, pl[j]);
    c[i + 1] += c[i] >> 29;
  }
#pragma unroll
  for (int k = 9; k < 18; k++) {
    r.l[k - 9] = opaque29((uint32_t)c[k] & M29);
    c[k + 1] += c[k] >> 29;
  }
  return r;
}

// ---- packing: 8 x 32-bit words <-> 9 x 29-bit limbs (value unchanged) ----
template <class C>
FB_HD Fl<C> unpack29(const uint32_t* w) {
  Fl<C> r;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const int bit = 29 * i, word = bit >> 5, sh = bit & 31;
    uint32_t v = w[word] >> sh;
    if (sh > 3 && word + 1 < 8) v |= w[word + 1] << (32 - sh);
    r.l[i] = v & M29;
  }
  return r;
}
template <class C>
FB_HD void pack29(const Fl<C>& a, uint32_t* w) {  // a must be canonical (< p < 2^254)
#pragma unroll
  for (int j = 0; j < 8; j++) {
    // word j covers bits [32j, 32j+32)
    const int lo_limb = (32 * j) / 29, sh = 32 * j - 29 * lo_limb;
    uint32_t v = a.l[lo_limb] >> sh;
    if (lo_limb + 1 < 9) v |= a.l[lo_limb + 1] << (29 - sh);
    if (29 - sh + 29 < 32 && lo_limb + 2 < 9) v |= a.l[lo_limb + 2] << (58 - sh);
    w[j] = v;
  }
}

// weakly reduced -> canonical [0, p)
template <class C>
FB_HD Fl<C> canon(const Fl<C>& a) {
  Fl<C> d;
  int32_t c = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    int32_t t = (int32_t)a.l[i] - (int32_t)C::p(i) + c;
    d.l[i] = (uint32_t)t & M29;
    c = t >> 29;
  }
  Fl<C> r;
#pragma unroll
  for (int i = 0; i < 9; i++) r.l[i] = c < 0 ? a.l[i] : d.l[i];
  return r;
}

// internal form (x * 2^261) -> the reference's memory form (x * 2^256), 8 x u32 canonical
template <class C>
FB_HD Fp<typename C::Base> to_standard(const Fl<C>& a) {
  Fl<C> k;
#pragma unroll
  for (int i = 0; i < 9; i++) k.l[i] = C::to_std(i);
  Fl<C> s = canon(mul(a, k));
  Fp<typename C::Base> r;
  pack29(s, r.v);
  return r;
}
// memory form already holding the internal value (bases converted once at key load)
template <class C>
FB_HD Fl<C> load_internal(const Fp<typename C::Base>& w) { return unpack29<C>(w.v); }
// standard Montgomery (x * 2^256) -> packed internal value (x * 2^261), done with the 32-bit multiplier
template <class B>
FB_HD Fp<B> std_to_internal_words(const Fp<B>& x, const uint32_t* c) {
  Fp<B> k;
  for (int i = 0; i < 8; i++) k.v[i] = c[i];
  return mul(x, k);
}

using Fq29 = Fl<Fq29Cfg>;
using Fr29 = Fl<Fr29Cfg>;

// ---- Fq2 over the lazy base field ----
struct Fq2L {
  Fq29 c0, c1;
  FB_HD static Fq2L zero() { return {Fq29::zero(), Fq29::zero()}; }
  FB_HD static Fq2L one() { return {Fq29::one(), Fq29::zero()}; }
  FB_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
};
FB_HD Fq2L add(const Fq2L& a, const Fq2L& b) { return {add(a.c0, b.c0), add(a.c1, b.c1)}; }
FB_HD Fq2L sub(const Fq2L& a, const Fq2L& b) { return {sub(a.c0, b.c0), sub(a.c1, b.c1)}; }
FB_HD Fq2L dbl(const Fq2L& a) { return {dbl(a.c0), dbl(a.c1)}; }
FB_HD Fq2L neg(const Fq2L& a) { return {neg(a.c0), neg(a.c1)}; }
FB_HD Fq2L mul(const Fq2L& a, const Fq2L& b) {
  Fq29 t0 = mul(a.c0, b.c0);
  Fq29 t1 = mul(a.c1, b.c1);
  Fq29 t2 = mul(add(a.c0, a.c1), add(b.c0, b.c1));
  return {sub(t0, t1), sub(sub(t2, t0), t1)};
}
FB_HD Fq2L sqr(const Fq2L& a) {
  Fq29 t = mul(a.c0, a.c1);
  return {mul(add(a.c0, a.c1), sub(a.c0, a.c1)), dbl(t)};
}

}  // namespace fb
