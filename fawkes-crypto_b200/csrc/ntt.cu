// Evaluation-domain transforms for the Groth16 H polynomial, BN254 Fr, sm_100a.
//
// Replaces bellman_ce's EvaluationDomain::{ifft,coset_fft,icoset_fft,mul_assign,
// sub_assign,divide_by_z_on_coset} (un-vendored crate; reached from
// fawkes-crypto/src/backend/bellman_groth16/prover.rs:80).  Conventions restated in
// m = next pow2 >= rows, omega = ROOT_OF_UNITY^(2^(28-k)),
// coset generator g = 7, H = icoset_fft((coset_fft(ifft a) * coset_fft(ifft b)
// - coset_fft(ifft c)) / (g^m - 1)), last coefficient dropped.
//
// B200 design (not bellman's radix-2 in-place loop):
//  * No bit-reversal passes.  Inverse transforms run decimation-in-frequency
//    (natural in, bit-reversed out), forward transforms decimation-in-time
//    (bit-reversed in, natural out); H comes out bit-reversed and the H bases are
//    stored bit-reversed once at key load.
//  * No scaling passes.  The coset shift is folded into the twiddles
//    (stage table = (g*w^j)^(2^t)), and all 1/m factors into the two constants of
//    the fused pointwise step  a*b*k1 - c*k2.
//  * Multi-pass shared-memory tiles (2^LOG_TILE elements, two conflict-free uint4
//    planes).  The last log2(tile) inverse stages and the first log2(tile) forward
//    stages share one contiguous tile ("mid" pass), so ifft+coset_fft costs
//    2*S+1 trips through HBM instead of 2*S+2.
//  * The pointwise (a*b-c)/Z step is fused into the first pass of the final
//    inverse transform.
#include "ntt.cuh"
#include "internal.h"

#include <cstdio>

namespace fb {

// --------------------------------------------------------------- tables ---
// tab[off_t + j] = base_j^(2^t), base_j = shift * root^j, j < 2^(k-1-t), off_t = 2^k - 2^(k-t)
__global__ void k_gen_twiddles(Fr* tab, int k, Fr root, Fr shift) {
  const uint64_t total = (1ull << k) - 1;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  // root^(2^i)
  __shared__ Fr rp[32];
  if (threadIdx.x == 0) {
    Fr x = root;
    for (int i = 0; i < k; i++) { rp[i] = x; x = sqr(x); }
  }
  __syncthreads();
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    // find stage t with off_t <= e < off_{t+1}:  off_t = 2^k - 2^(k-t)
    uint64_t rem = (1ull << k) - e;  // in (2^(k-t-1), 2^(k-t)]
    int t = k - (64 - __clzll(rem - 1));  // rem-1 in [2^(k-t-1), 2^(k-t)) -> bitlen = k-t
    if (rem == 1) t = k - 1;              // e = 2^k - 1 - ... handled: rem-1 = 0
    uint64_t off = (1ull << k) - (1ull << (k - t));
    uint32_t j = (uint32_t)(e - off);
    Fr x = shift;
    for (int i = 0; i < k; i++)
      if ((j >> i) & 1) x = mul(x, rp[i]);
    for (int i = 0; i < t; i++) x = sqr(x);
    tab[e] = x;
  }
}

// ------------------------------------------------------- shared-mem tile ---
struct Tile {
  uint4* lo;
  uint4* hi;
  __device__ __forceinline__ Fr get(int i) const {
    Fr r;
    uint4 a = lo[i], b = hi[i];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
  }
  __device__ __forceinline__ void put(int i, const Fr& r) const {
    lo[i] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    hi[i] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
  }
};

__device__ __forceinline__ Fr ldg_fr(const Fr* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ Fr ld_fr(const Fr* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void st_fr(Fr* p, const Fr& r) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// twiddle for table-stage t, index j.  NEGINV: w^-e = (j==0) ? 1 : -tab_t[len-j]
template <bool NEGINV>
__device__ __forceinline__ Fr twiddle(const Fr* tab, int k, int t, uint32_t j) {
  const uint64_t off = (1ull << k) - (1ull << (k - t));
  if (!NEGINV) return ldg_fr(tab + off + j);
  if (j == 0) return Fr::one();
  const uint32_t len = 1u << (k - 1 - t);
  return neg(ldg_fr(tab + off + (len - j)));
}

// DIF stages over the tile's b "mid" bits (largest half first)
// k: log2 of the GLOBAL transform (twiddle tables); lb: bit position of the tile's mid bits in the
// LOCAL array; sh: distributed layout, global index = (local << sh.shift) | sh.low (shift = 0 on one GPU)
struct NttShard { int shift; uint32_t low; };
template <bool NEGINV>
__device__ __forceinline__ void dif_stages(const Tile& s, const Fr* tab, int k, int lb, int a,
                                           int b, uint32_t lo0, NttShard sh) {
  const int nb = 1 << (a + b - 1);
  for (int d = 0; d < b; d++) {
    const int hm = 1 << (b - 1 - d);
    const int t = k - (lb + sh.shift) - b + d;
    for (int q = threadIdx.x; q < nb; q += blockDim.x) {
      const int l = q & ((1 << a) - 1);
      const int qm = q >> a;
      const int jm = qm & (hm - 1);
      const int mid = ((qm - jm) << 1) + jm;
      const int i0 = (mid << a) + l, i1 = ((mid + hm) << a) + l;
      const uint32_t j = ((((uint32_t)jm << lb) + lo0 + l) << sh.shift) | sh.low;
      Fr w = twiddle<NEGINV>(tab, k, t, j);
      Fr u = s.get(i0), v = s.get(i1);
      s.put(i0, add(u, v));
      s.put(i1, mul(sub(u, v), w));
    }
    __syncthreads();
  }
}
// DIT stages over the tile's b "mid" bits (smallest half first)
template <bool NEGINV>
__device__ __forceinline__ void dit_stages(const Tile& s, const Fr* tab, int k, int lb, int a,
                                           int b, uint32_t lo0, NttShard sh) {
  const int nb = 1 << (a + b - 1);
  for (int d = 0; d < b; d++) {
    const int hm = 1 << d;
    const int t = k - 1 - (lb + sh.shift + d);
    for (int q = threadIdx.x; q < nb; q += blockDim.x) {
      const int l = q & ((1 << a) - 1);
      const int qm = q >> a;
      const int jm = qm & (hm - 1);
      const int mid = ((qm - jm) << 1) + jm;
      const int i0 = (mid << a) + l, i1 = ((mid + hm) << a) + l;
      const uint32_t j = ((((uint32_t)jm << lb) + lo0 + l) << sh.shift) | sh.low;
      Fr w = twiddle<NEGINV>(tab, k, t, j);
      Fr u = s.get(i0), v = mul(s.get(i1), w);
      s.put(i0, add(u, v));
      s.put(i1, sub(u, v));
    }
    __syncthreads();
  }
}

// One pass over a tile of 2^b (mid bits [lb,lb+b)) x 2^a (consecutive lo) elements.
// MODE 0: DIF stages (tab0, NEG0)         MODE 1: DIT stages (tab0, NEG0)
// MODE 2: contiguous mid pass: DIF with tab0 (negated-inverse plain) then DIT with tab1
// MODE 3: DIF stages, input = x*y*k1 - z*k2 (pointwise fused), written to x
template <int MODE, bool NEG0>
__global__ void __launch_bounds__(NTT_THREADS)
k_ntt_pass(Fr* x, const Fr* y, const Fr* z, const Fr* tab0, const Fr* tab1, int k, int lb, int a,
           int b, Fr k1, Fr k2, NttShard sh, uint64_t bstride) {
  extern __shared__ uint4 smem[];
  // batched transforms (fb_prove_batch): blockIdx.y selects one of gridDim.y arrays, bstride elements apart
  x += (uint64_t)blockIdx.y * bstride;
  if (MODE == 3) { y += (uint64_t)blockIdx.y * bstride; z += (uint64_t)blockIdx.y * bstride; }
  Tile s{smem, smem + (1 << (a + b))};
  const int tile = 1 << (a + b);
  // block -> (hi, lo chunk)
  const uint32_t chunks = 1u << (lb - a);
  const uint32_t hi = blockIdx.x / chunks;
  const uint32_t lo0 = (blockIdx.x % chunks) << a;
  const uint64_t base = ((uint64_t)hi << (lb + b)) + lo0;
  for (int e = threadIdx.x; e < tile; e += blockDim.x) {
    const int l = e & ((1 << a) - 1), mid = e >> a;
    const uint64_t g = base + ((uint64_t)mid << lb) + l;
    Fr v = ld_fr(x + g);
    if (MODE == 3) {
      Fr vy = ld_fr(y + g), vz = ld_fr(z + g);
      v = sub(mul(mul(v, vy), k1), mul(vz, k2));
    }
    s.put(e, v);
  }
  __syncthreads();
  if (MODE == 0 || MODE == 3) dif_stages<NEG0>(s, tab0, k, lb, a, b, lo0, sh);
  if (MODE == 1) dit_stages<NEG0>(s, tab0, k, lb, a, b, lo0, sh);
  if (MODE == 2) {
    dif_stages<true>(s, tab0, k, lb, a, b, lo0, sh);
    dit_stages<false>(s, tab1, k, lb, a, b, lo0, sh);
  }
  for (int e = threadIdx.x; e < tile; e += blockDim.x) {
    const int l = e & ((1 << a) - 1), mid = e >> a;
    const uint64_t g = base + ((uint64_t)mid << lb) + l;
    st_fr(x + g, s.get(e));
  }
}

__global__ void k_bitrev_permute(Fr* dst, const Fr* src, int k) {
  const uint64_t n = 1ull << k;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t r = __brev((uint32_t)i) >> (32 - k);
    st_fr(dst + r, ld_fr(src + i));
  }
}

__global__ void k_scale(Fr* x, uint64_t n, Fr s) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    st_fr(x + i, mul(ld_fr(x + i), s));
}

// ----------------------------------------------------------------- host ---
static Fr host_pow(Fr a, uint64_t e) { return pow_u64(a, e); }

int NttDomain::init(int k_, cudaStream_t st) {
  k = k_;
  if (k < 1 || k >= 28) return -1;  // bellman: exp >= Fr::S is PolynomialDegreeTooLarge
  // omega = ROOT_OF_UNITY^(2^(28-k))
  Fr root;
  {
    constexpr uint32_t v[8] = {0xb639feb8u, 0x9632c7c5u, 0x0d0ff299u, 0x985ce340u,
                               0x01b0ecd8u, 0xb2dd8800u, 0x6d98ce29u, 0x1d69070du};
    for (int i = 0; i < 8; i++) root.v[i] = v[i];
  }
  for (int i = k; i < 28; i++) root = sqr(root);
  omega = root;
  Fr g;
  {
    constexpr uint32_t v[8] = {0x4fffffdbu, 0x3057819eu, 0x6832bb01u, 0x307f6d86u,
                               0x484e3a89u, 0x5c65ec9fu, 0x73d3d9f8u, 0x0180a965u};
    for (int i = 0; i < 8; i++) g.v[i] = v[i];
  }
  const uint64_t m = 1ull << k;
  Fr omega_inv = inv(omega), g_inv = inv(g);
  Fr mm = Fr::zero();
  mm.v[0] = (uint32_t)m;
  mm = to_mont(mm);
  Fr minv = inv(mm);
  Fr zinv = inv(sub(host_pow(g, m), Fr::one()));  // 1/(g^m - 1)
  // pointwise constants: a'=m*a etc. and the final transform also owes 1/m
  Fr m2 = mul(minv, minv);
  k2 = mul(zinv, m2);          // zinv / m^2
  k1 = mul(k2, minv);          // zinv / m^3
  this->minv = minv;
  size_t bytes = (m - 1) * sizeof(Fr);
  if (cudaMalloc(&tab_plain, bytes) != cudaSuccess) return -2;
  if (cudaMalloc(&tab_coset, bytes) != cudaSuccess) return -2;
  if (cudaMalloc(&tab_icoset, bytes) != cudaSuccess) return -2;
  int blocks = (int)std::min<uint64_t>((m + 255) / 256, 148 * 8);
  k_gen_twiddles<<<blocks, 256, 0, st>>>(tab_plain, k, omega, Fr::one());
  k_gen_twiddles<<<blocks, 256, 0, st>>>(tab_coset, k, omega, g);
  k_gen_twiddles<<<blocks, 256, 0, st>>>(tab_icoset, k, omega_inv, g_inv);
  // pass plan
  plan_bm = std::min(k, NTT_LOG_TILE);
  int rem = k - plan_bm;
  n_strided = 0;
  if (rem > 0) {
    int maxb = NTT_LOG_TILE - NTT_MIN_LO;
    n_strided = (rem + maxb - 1) / maxb;
    int lb = plan_bm;
    for (int i = 0; i < n_strided; i++) {
      int b = rem / n_strided + (i < rem % n_strided ? 1 : 0);
      pass_lb[i] = lb;
      pass_b[i] = b;
      lb += b;
    }
  }
  static bool attr_done = false;
  if (!attr_done) {
    size_t sm = (size_t)sizeof(Fr) << NTT_LOG_TILE;
    cudaFuncSetAttribute(k_ntt_pass<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(k_ntt_pass<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(k_ntt_pass<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(k_ntt_pass<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(k_ntt_pass<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(k_ntt_pass<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    attr_done = true;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

void NttDomain::destroy() {
  cudaFree(tab_plain);
  cudaFree(tab_coset);
  cudaFree(tab_icoset);
  tab_plain = tab_coset = tab_icoset = nullptr;
}

// kl = log2 of the local array length (= k on one GPU)
template <int MODE, bool NEG0>
static void launch_pass(Fr* x, const Fr* y, const Fr* z, const Fr* t0, const Fr* t1, int k, int lb,
                        int b, const Fr& k1, const Fr& k2, cudaStream_t st, int kl = -1,
                        NttShard sh = NttShard{0, 0}, unsigned batch = 1, uint64_t bstride = 0) {
  if (kl < 0) kl = k;
  int a = std::min(lb, NTT_LOG_TILE - b);
  size_t sm = (size_t)sizeof(Fr) << (a + b);
  unsigned blocks = 1u << (kl - a - b);
  kstat_begin(KSTAT_NTT, st);
  k_ntt_pass<MODE, NEG0><<<dim3(blocks, batch), NTT_THREADS, sm, st>>>(x, y, z, t0, t1, k, lb, a, b, k1, k2, sh, bstride);
  kstat_end(KSTAT_NTT, st);
  count_launch();
}

// x (natural order evaluations on H) -> natural order evaluations on gH, both scaled by m
void NttDomain::ifft_then_coset_fft(Fr* x, cudaStream_t st, unsigned batch, uint64_t bstride) const {
  Fr z = Fr::zero();
  const NttShard one{0, 0};
  for (int i = n_strided - 1; i >= 0; i--)
    launch_pass<0, true>(x, nullptr, nullptr, tab_plain, nullptr, k, pass_lb[i], pass_b[i], z, z, st, -1, one, batch, bstride);
  launch_pass<2, true>(x, nullptr, nullptr, tab_plain, tab_coset, k, 0, plan_bm, z, z, st, -1, one, batch, bstride);
  for (int i = 0; i < n_strided; i++)
    launch_pass<1, false>(x, nullptr, nullptr, tab_coset, nullptr, k, pass_lb[i], pass_b[i], z, z, st, -1, one, batch, bstride);
}

// a <- icoset_fft((a*b - c)/Z) in BIT-REVERSED coefficient order, exact (1/m folded)
void NttDomain::pointwise_then_icoset_fft(Fr* a, const Fr* b, const Fr* c, cudaStream_t st, unsigned batch,
                                          uint64_t bstride) const {
  Fr z = Fr::zero();
  const NttShard one{0, 0};
  if (n_strided == 0) {
    launch_pass<3, false>(a, b, c, tab_icoset, nullptr, k, 0, plan_bm, k1, k2, st, -1, one, batch, bstride);
    return;
  }
  for (int i = n_strided - 1; i >= 0; i--) {
    if (i == n_strided - 1)
      launch_pass<3, false>(a, b, c, tab_icoset, nullptr, k, pass_lb[i], pass_b[i], k1, k2, st, -1, one, batch, bstride);
    else
      launch_pass<0, false>(a, nullptr, nullptr, tab_icoset, nullptr, k, pass_lb[i], pass_b[i], z, z, st, -1, one, batch, bstride);
  }
  launch_pass<0, false>(a, nullptr, nullptr, tab_icoset, nullptr, k, 0, plan_bm, z, z, st, -1, one, batch, bstride);
}

// Stand-alone transforms for tests / setup.  Natural order in and out.
// kind 0: fft (w)   1: ifft (w^-1, /m)   2: coset_fft   3: icoset_fft
void NttDomain::transform(Fr* x, Fr* scratch, int kind, cudaStream_t st) const {
  Fr z = Fr::zero();
  unsigned pb = (unsigned)std::min<uint64_t>(((1ull << k) + 255) / 256, 148 * 16);
  if (kind == 0 || kind == 2) {  // forward: bitrev -> DIT
    const Fr* tab = kind == 0 ? tab_plain : tab_coset;
    k_bitrev_permute<<<pb, 256, 0, st>>>(scratch, x, k);
    launch_pass<1, false>(scratch, nullptr, nullptr, tab, nullptr, k, 0, plan_bm, z, z, st);
    for (int i = 0; i < n_strided; i++)
      launch_pass<1, false>(scratch, nullptr, nullptr, tab, nullptr, k, pass_lb[i], pass_b[i], z, z, st);
    cudaMemcpyAsync(x, scratch, sizeof(Fr) << k, cudaMemcpyDeviceToDevice, st);
  } else {
    for (int i = n_strided - 1; i >= 0; i--) {
      if (kind == 1)
        launch_pass<0, true>(x, nullptr, nullptr, tab_plain, nullptr, k, pass_lb[i], pass_b[i], z, z, st);
      else
        launch_pass<0, false>(x, nullptr, nullptr, tab_icoset, nullptr, k, pass_lb[i], pass_b[i], z, z, st);
    }
    if (kind == 1)
      launch_pass<0, true>(x, nullptr, nullptr, tab_plain, nullptr, k, 0, plan_bm, z, z, st);
    else
      launch_pass<0, false>(x, nullptr, nullptr, tab_icoset, nullptr, k, 0, plan_bm, z, z, st);
    k_scale<<<pb, 256, 0, st>>>(x, 1ull << k, minv);
    k_bitrev_permute<<<pb, 256, 0, st>>>(scratch, x, k);
    cudaMemcpyAsync(x, scratch, sizeof(Fr) << k, cudaMemcpyDeviceToDevice, st);
  }
}

// ------------------------------------------------------------ distributed ---
// G = 2^g ranks.  "cyclic" layout: rank r holds the elements with global index = r (mod G), local
// index = global >> g; every butterfly on global bits >= g is local.  "block" layout: rank r holds the
// contiguous range [r * m/G, (r+1) * m/G); every butterfly on global bits < k - g is local.
// ifft + coset_fft:  cyclic DIF on bits [bm, k)  -> exchange ->  block mid pass on bits [0, bm)
//                    -> exchange ->  cyclic DIT on bits [bm, k).          (bm = 12, g <= bm <= k - g)
// final:             cyclic pointwise + DIF on bits [bm, k) -> exchange -> block DIF on bits [0, bm):
//                    H comes out bit-reversed in block layout = the rank's contiguous H-base shard.
// One exchange = an all-to-all of m/G elements per rank; a, b, c travel in one grouped call.
__global__ void k_unpack_to_block(Fr* __restrict__ dst, const Fr* __restrict__ src, int g, uint64_t C) {
  // src[r * C + q] (chunk received from rank r)  ->  dst[q * G + r]
  const uint64_t n = C << g;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t q = i >> g, r = i & ((1u << g) - 1);
    st_fr(dst + i, ld_fr(src + r * C + q));
  }
}
__global__ void k_pack_from_block(Fr* __restrict__ dst, const Fr* __restrict__ src, int g, uint64_t C) {
  // src[q * G + r]  ->  dst[r * C + q] (chunk to send to rank r)
  const uint64_t n = C << g;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t q = i >> g, r = i & ((1u << g) - 1);
    st_fr(dst + r * C + q, ld_fr(src + i));
  }
}

bool NttDomain::dist_supported(int g) const {
  return g >= 1 && k > NTT_LOG_TILE && k - g >= NTT_LOG_TILE && NTT_LOG_TILE - g >= NTT_MIN_LO;
}

void NttDomain::dist_plan(int g, int* n_pass, int* lb, int* b) const {
  // strided cyclic passes over global bits [bm, k) = local bits [bm - g, k - g)
  const int bm = NTT_LOG_TILE, rem = k - bm, maxb = NTT_LOG_TILE - NTT_MIN_LO;
  const int np = (rem + maxb - 1) / maxb;
  int pos = bm - g;
  for (int i = 0; i < np; i++) {
    b[i] = rem / np + (i < rem % np ? 1 : 0);
    lb[i] = pos;
    pos += b[i];
  }
  *n_pass = np;
}

static unsigned grid_for(uint64_t n) { return (unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 16); }

int NttDomain::dist_h_pipeline(Fr* const ev[3], Fr* const tmp[3], int g, int rank, NttExchange* xch,
                               cudaStream_t st) const {
  const int kl = k - g, bm = NTT_LOG_TILE;
  const uint64_t ml = 1ull << kl, C = ml >> g;
  int np, lb[8], b[8];
  dist_plan(g, &np, lb, b);
  const NttShard cyc{g, (uint32_t)rank}, blk{0, 0};
  Fr z = Fr::zero();
  cudaStream_t cs = xch->side_stream();
  int rc = 0;
  if (cs) {
    // Overlapped schedule: the exchange of one array runs on the side stream while the compute stream transforms
    // the next one (every rank issues the exchanges in the same order: a, b, c, then a, b, c).
    auto ev_ = [&](int i) { return xch->event(i); };
    // 1 + 2. cyclic inverse DIF on the high bits, then cyclic -> block, array by array
    for (int v = 0; v < 3; v++) {
      for (int i = np - 1; i >= 0; i--)
        launch_pass<0, true>(ev[v], nullptr, nullptr, tab_plain, nullptr, k, lb[i], b[i], z, z, st, kl, cyc);
      cudaEventRecord(ev_(v), st);
      cudaStreamWaitEvent(cs, ev_(v), 0);
      const Fr* send1[1] = {ev[v]};
      Fr* recv1[1] = {tmp[v]};
      kstat_begin(KSTAT_EXCHANGE, cs);
      rc = xch->all_to_all(send1, recv1, 1, C, cs);
      if (rc) return rc;
      kstat_end(KSTAT_EXCHANGE, cs);
      cudaEventRecord(ev_(3 + v), cs);
    }
    // 3 + 4. block layout: last bm inverse stages + first bm forward (coset) stages in one tile, then block -> cyclic
    for (int v = 0; v < 3; v++) {
      cudaStreamWaitEvent(st, ev_(3 + v), 0);
      k_unpack_to_block<<<grid_for(ml), 256, 0, st>>>(ev[v], tmp[v], g, C);
      launch_pass<2, true>(ev[v], nullptr, nullptr, tab_plain, tab_coset, k, 0, bm, z, z, st, kl, blk);
      k_pack_from_block<<<grid_for(ml), 256, 0, st>>>(tmp[v], ev[v], g, C);
      cudaEventRecord(ev_(6 + v), st);
      cudaStreamWaitEvent(cs, ev_(6 + v), 0);
      const Fr* send1[1] = {tmp[v]};
      Fr* recv1[1] = {ev[v]};
      kstat_begin(KSTAT_EXCHANGE, cs);
      rc = xch->all_to_all(send1, recv1, 1, C, cs);
      if (rc) return rc;
      kstat_end(KSTAT_EXCHANGE, cs);
      cudaEventRecord(ev_(9 + v), cs);
    }
    // 5. cyclic forward (coset) DIT on the high bits
    for (int v = 0; v < 3; v++) {
      cudaStreamWaitEvent(st, ev_(9 + v), 0);
      for (int i = 0; i < np; i++)
        launch_pass<1, false>(ev[v], nullptr, nullptr, tab_coset, nullptr, k, lb[i], b[i], z, z, st, kl, cyc);
    }
  } else {
  // 1. cyclic inverse DIF on the high bits (three arrays)
  for (int v = 0; v < 3; v++)
    for (int i = np - 1; i >= 0; i--)
      launch_pass<0, true>(ev[v], nullptr, nullptr, tab_plain, nullptr, k, lb[i], b[i], z, z, st, kl, cyc);
  // 2. cyclic -> block: contiguous chunks out, strided placement in
  const Fr* send3[3] = {ev[0], ev[1], ev[2]};
  kstat_begin(KSTAT_EXCHANGE, st);
  rc = xch->all_to_all(send3, tmp, 3, C, st);
  if (rc) return rc;
  for (int v = 0; v < 3; v++) k_unpack_to_block<<<grid_for(ml), 256, 0, st>>>(ev[v], tmp[v], g, C);
  kstat_end(KSTAT_EXCHANGE, st);
  // 3. block layout: last bm inverse stages + first bm forward (coset) stages in one tile
  for (int v = 0; v < 3; v++)
    launch_pass<2, true>(ev[v], nullptr, nullptr, tab_plain, tab_coset, k, 0, bm, z, z, st, kl, blk);
  // 4. block -> cyclic
  kstat_begin(KSTAT_EXCHANGE, st);
  for (int v = 0; v < 3; v++) k_pack_from_block<<<grid_for(ml), 256, 0, st>>>(tmp[v], ev[v], g, C);
  const Fr* send3b[3] = {tmp[0], tmp[1], tmp[2]};
  rc = xch->all_to_all(send3b, ev, 3, C, st);
  if (rc) return rc;
  kstat_end(KSTAT_EXCHANGE, st);
  // 5. cyclic forward (coset) DIT on the high bits
  for (int v = 0; v < 3; v++)
    for (int i = 0; i < np; i++)
      launch_pass<1, false>(ev[v], nullptr, nullptr, tab_coset, nullptr, k, lb[i], b[i], z, z, st, kl, cyc);
  }
  // 6. pointwise + cyclic inverse-coset DIF on the high bits
  for (int i = np - 1; i >= 0; i--) {
    if (i == np - 1)
      launch_pass<3, false>(ev[0], ev[1], ev[2], tab_icoset, nullptr, k, lb[i], b[i], k1, k2, st, kl, cyc);
    else
      launch_pass<0, false>(ev[0], nullptr, nullptr, tab_icoset, nullptr, k, lb[i], b[i], z, z, st, kl, cyc);
  }
  // 7. cyclic -> block, then the low bm stages: H in bit-reversed order, block layout
  const Fr* send1[1] = {ev[0]};
  Fr* recv1[1] = {tmp[0]};
  kstat_begin(KSTAT_EXCHANGE, st);
  rc = xch->all_to_all(send1, recv1, 1, C, st);
  if (rc) return rc;
  k_unpack_to_block<<<grid_for(ml), 256, 0, st>>>(ev[0], tmp[0], g, C);
  kstat_end(KSTAT_EXCHANGE, st);
  launch_pass<0, false>(ev[0], nullptr, nullptr, tab_icoset, nullptr, k, 0, bm, z, z, st, kl, blk);
  count_launch(10);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

void NttDomain::bitrev(Fr* dst, const Fr* src, cudaStream_t st) const {
  unsigned pb = (unsigned)std::min<uint64_t>(((1ull << k) + 255) / 256, 148 * 16);
  k_bitrev_permute<<<pb, 256, 0, st>>>(dst, src, k);
}

}  // namespace fb
