// BN254 Fq / Fr Montgomery arithmetic on 5 x 52-bit limbs, multiplied on the FP64 pipe.
//
// Why: on sm_100a (B200) IMAD.WIDE.U32 issues once per 4 cycles per SM sub-partition, so the 32-bit
// limb multiplier of ff.cuh is bound at 136 x 4 = 544 cycles per multiply -- while DFMA issues once per
// 2 cycles on its own pipe (64 FP64 lanes per SM; tools/pipe_probe.cu, profiles/r01_pipe_probe_*.txt).
// A double holds a 52-bit limb exactly, and two fused multiply-adds split a 52 x 52 product exactly:
//     h = fma_rz(a, b, 2^104)            = 2^104 + 2^52 * floor(ab / 2^52)
//     l = fma_rz(a, b, 2^104 + 2^52 - h) = 2^52 + (ab mod 2^52)
// (the rounding-toward-zero of the first is the floor; the second is exact).  The bit patterns of h and l
// are (exponent | 52-bit integer), so they are summed per column as 64-bit INTEGERS on the ALU pipe and
// the exponent words are cancelled by constants folded into the column initialisers.  A Montgomery product
// is 25 + 5 + 25 limb products = ~175 FP64 instructions + ~170 ALU instructions and NO wide integer MAC,
// so warps running this multiplier and warps running the IMAD one share an SM without competing for a pipe.
//
// Representation: value = sum l[i] 2^(52 i), limbs < 2^52 (normalised), Montgomery radix 2^260.  Because
// 2^260 > 84 p, a product of operands below ~9 p is below 2 p with no conditional subtraction, which lets
// the curve formulas use plain "a - b + k p" subtractions.  Relation to the reference's memory form
// (x 2^256, ff-uint/src/num/mod.rs:21-23): load with to52<4> (16 x 2^256 = x 2^260), store with
// from52_std (one multiply by 2^256 mod p), results are unique field elements, hence bit-identical.
#pragma once
#include "ff.cuh"

namespace fb {

struct Fq52Cfg {
  using Base = FqCfg;
  static constexpr uint64_t NP = 0x20782e4866389ull;  // -p^-1 mod 2^52
  FB_HD static constexpr uint64_t p(int i) {
    constexpr uint64_t v[5] = {0x8c16d87cfd47ull, 0x916871ca8d3c2ull, 0x181585d97816aull, 0xa029b85045b68ull,
                               0x30644e72e131ull};
    return v[i];
  }
};
struct Fr52Cfg {
  using Base = FrCfg;
  static constexpr uint64_t NP = 0x1f593efffffffull;
  FB_HD static constexpr uint64_t p(int i) {
    constexpr uint64_t v[5] = {0x1f593f0000001ull, 0x4879b9709143eull, 0x181585d2833e8ull, 0xa029b85045b68ull,
                               0x30644e72e131ull};
    return v[i];
  }
};

template <class C>
struct F52 {
  uint64_t l[5];
  using Cfg = C;
};

#if defined(__CUDA_ARCH__)
namespace f52 {
constexpr uint64_t MASK = (1ull << 52) - 1;
constexpr uint64_t LO_BIAS = 0x4330000000000000ull;  // bits of 2^52
constexpr uint64_t HI_BIAS = 0x4670000000000000ull;  // bits of 2^104
// how many lo / hi bit patterns column k receives over the whole multiply (product + reduction rows);
// the quotient rows skip lo(q p_0): that limb is known to cancel, only its carry is needed
__device__ __forceinline__ constexpr int n_lo(int k) {
  int c = 0;
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < 5; j++) c += (i + j == k);        // product
  for (int r = 0; r < 5; r++)
    for (int j = 1; j < 5; j++) c += (r + j == k);        // reduction
  return c;
}
__device__ __forceinline__ constexpr int n_hi(int k) {
  int c = 0;
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < 5; j++) c += (i + j + 1 == k);
  for (int r = 0; r < 5; r++)
    for (int j = 0; j < 5; j++) c += (r + j + 1 == k);
  return c;
}
__device__ __forceinline__ constexpr uint64_t col_init(int k) {
  return 0ull - ((uint64_t)n_lo(k) * LO_BIAS + (uint64_t)n_hi(k) * HI_BIAS);
}
__device__ __forceinline__ double to_double(uint64_t limb) {  // exact for limb < 2^52
  return __longlong_as_double((long long)(limb | LO_BIAS)) - 0x1p52;
}
// col[hi_k] += bits(h), col[lo_k] += bits(l) for the exact split of a * b
__device__ __forceinline__ void mac_split(double a, double b, uint64_t& lo_col, uint64_t& hi_col) {
  const double h = __fma_rz(a, b, 0x1p104);
  const double s = (0x1p104 + 0x1p52) - h;
  const double l = __fma_rz(a, b, s);
  hi_col += (uint64_t)__double_as_longlong(h);
  lo_col += (uint64_t)__double_as_longlong(l);
}
}  // namespace f52

// a * b * 2^-260 mod p, result < a b / 2^260 + p, limbs normalised
template <class C>
__device__ __forceinline__ F52<C> mul52(const F52<C>& a, const F52<C>& b) {
  using namespace f52;
  double ad[5], bd[5];
#pragma unroll
  for (int i = 0; i < 5; i++) { ad[i] = to_double(a.l[i]); bd[i] = to_double(b.l[i]); }
  uint64_t col[10];
#pragma unroll
  for (int k = 0; k < 10; k++) col[k] = col_init(k);
#pragma unroll
  for (int i = 0; i < 5; i++)
#pragma unroll
    for (int j = 0; j < 5; j++) mac_split(ad[i], bd[j], col[i + j], col[i + j + 1]);
#pragma unroll
  for (int r = 0; r < 5; r++) {
    const uint64_t t = col[r] & MASK;
    const double td = to_double(t);
    // q = t * (-p^-1) mod 2^52
    const double hq = __fma_rz(td, (double)C::NP, 0x1p104);
    const double sq = (0x1p104 + 0x1p52) - hq;
    const double qd = __fma_rz(td, (double)C::NP, sq) - 0x1p52;
    // col[r] + lo(q p_0) == 0 mod 2^52: only its carry is needed
    col[r + 1] += (uint64_t)__double_as_longlong(__fma_rz(qd, (double)C::p(0), 0x1p104));
#pragma unroll
    for (int j = 1; j < 5; j++) mac_split(qd, (double)C::p(j), col[r + j], col[r + j + 1]);
    col[r + 1] += (col[r] >> 52) + (t != 0 ? 1u : 0u);
  }
  F52<C> out;
  uint64_t cy = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint64_t v = col[5 + k] + cy;
    out.l[k] = v & MASK;
    cy = v >> 52;
  }
  out.l[4] = col[9] + cy;
  return out;
}

// 8 x 32-bit limbs (value V < 2^256) -> 52-bit limbs of V << SH (SH = 4 turns x 2^256 into x 2^260)
template <class C, int SH>
__device__ __forceinline__ F52<C> to52(const Fp<typename C::Base>& x) {
  uint64_t w[5];
#pragma unroll
  for (int k = 0; k < 4; k++) w[k] = (uint64_t)x.v[2 * k] | ((uint64_t)x.v[2 * k + 1] << 32);
  w[4] = 0;
  F52<C> r;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    const int pos = 52 * i - SH;
    uint64_t v;
    if (pos < 0) {
      v = w[0] << (-pos);
    } else {
      const int q = pos >> 6, s = pos & 63;
      v = w[q] >> s;
      if (s != 0 && q + 1 <= 4) v |= w[q + 1] << (64 - s);
    }
    r.l[i] = i < 4 ? (v & f52::MASK) : v;  // top limb keeps the SH extra bits
  }
  if (SH > 0) r.l[4] &= (1ull << (48 + SH)) - 1;
  return r;
}
// normalised limbs with value < 2^256 -> 8 x 32-bit limbs
template <class C>
__device__ __forceinline__ Fp<typename C::Base> from52_raw(const F52<C>& a) {
  uint64_t w[4];
  w[0] = a.l[0] | (a.l[1] << 52);
  w[1] = (a.l[1] >> 12) | (a.l[2] << 40);
  w[2] = (a.l[2] >> 24) | (a.l[3] << 28);
  w[3] = (a.l[3] >> 36) | (a.l[4] << 16);
  Fp<typename C::Base> r;
#pragma unroll
  for (int k = 0; k < 4; k++) { r.v[2 * k] = (uint32_t)w[k]; r.v[2 * k + 1] = (uint32_t)(w[k] >> 32); }
  return r;
}
// x 2^260 (any value below ~80 p) -> canonical x 2^256 in the reference's memory form
template <class C>
__device__ __forceinline__ Fp<typename C::Base> from52_std(const F52<C>& a) {
  const F52<C> one256 = to52<C, 0>(Fp<typename C::Base>::one());  // 2^256 mod p
  Fp<typename C::Base> r = from52_raw(mul52(a, one256));          // < a / 84 + p < 2 p
  ptx::cond_sub<typename C::Base>(r.v);
  return r;
}

// limbwise a + b, normalised (values stay far below 2^260 in the curve formulas)
template <class C>
__device__ __forceinline__ F52<C> add52(const F52<C>& a, const F52<C>& b) {
  F52<C> r;
  uint64_t cy = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint64_t v = a.l[i] + b.l[i] + cy;
    r.l[i] = v & f52::MASK;
    cy = v >> 52;
  }
  r.l[4] = a.l[4] + b.l[4] + cy;
  return r;
}
// a - b + K p, K p given as 52-bit limbs (compile-time multiple of p that exceeds any b passed in)
template <class C, int K>
__device__ __forceinline__ F52<C> sub52(const F52<C>& a, const F52<C>& b) {
  // limbs of K * p
  F52<C> r;
  int64_t cy = 0;
  unsigned __int128 kp = 0;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    // K * p limb i with carry from below, computed at compile time by constant folding
    kp += (unsigned __int128)C::p(i) * (unsigned)K;
    const uint64_t kl = i < 4 ? (uint64_t)(kp & f52::MASK) : (uint64_t)kp;
    kp >>= 52;
    const int64_t v = (int64_t)a.l[i] - (int64_t)b.l[i] + (int64_t)kl + cy;
    if (i < 4) {
      r.l[i] = (uint64_t)v & f52::MASK;
      cy = v >> 52;  // arithmetic
    } else {
      r.l[i] = (uint64_t)v;
    }
  }
  return r;
}
// Necessary condition for x == 0 mod p when x <= KMAX * p: the low limb equals that of some k * p.  False
// positives have probability ~KMAX * 2^-52; callers take an exact slow path when this fires.
template <class C, int KMAX>
__device__ __forceinline__ bool maybe_zero52(const F52<C>& x) {
  bool hit = false;
#pragma unroll
  for (int k = 0; k <= KMAX; k++) hit |= x.l[0] == ((C::p(0) * (uint64_t)k) & f52::MASK);
  return hit;
}
#endif  // __CUDA_ARCH__

using Fq52 = F52<Fq52Cfg>;
using Fr52 = F52<Fr52Cfg>;

}  // namespace fb
