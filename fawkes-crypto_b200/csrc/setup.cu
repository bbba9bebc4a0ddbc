// Groth16 parameter generation with an explicit trapdoor (harness-grade `setup`).
//
// Replaces bellman::groth16::generate_random_parameters as called from
// fawkes-crypto/src/backend/bellman_groth16/setup.rs:17-20 (restated in SURVEY.md
// App. C.4): same row set (circuit gates + one `input_i * 0 = 0` row per input), Lagrange
// values L_j(tau) by an inverse FFT of the powers of tau, h[i] = tau^i Z(tau)/delta * G1,
// a/b queries with the points at infinity filtered out, ic / l divided by gamma / delta.
// Generators are the standard ones ((1,2) and the EIP-197 G2 generator) instead of
// bellman's random g1, g2; output is bellman's Parameters byte format.
//
// B200 mapping: the Lagrange transform runs on the NTT kernels, every
// scalar -> point conversion on a fixed-base kernel (8-bit windows, 32 mixed adds per
// point, table resident in L2); the sparse column accumulation stays on the host.
#include "../../include/fawkes_b200.h"

#include <cstring>
#include <thread>

#include "host_fr.h"
#include "internal.h"

namespace fb {

static const uint32_t G2_GEN[4][8] = {
    // x.c0, x.c1, y.c0, y.c1 canonical little-endian limbs (EIP-197 generator)
    {0xd992f6edu, 0x46debd5cu, 0xf75edaddu, 0x674322d4u, 0x5e5c4479u, 0x426a0066u, 0x121f1e76u, 0x1800deefu},
    {0xaef312c2u, 0x97e485b7u, 0x35a9e712u, 0xf1aa4933u, 0x31fb5d25u, 0x7260bfb7u, 0x920d483au, 0x198e9393u},
    {0x66fa7daau, 0x4ce6cc01u, 0x0c43d37bu, 0xe3d1e769u, 0x8dcb408fu, 0x4aab7180u, 0xdb8c6debu, 0x12c85ea5u},
    {0xd122975bu, 0x55acdadcu, 0x70b38ef3u, 0xbc4b3133u, 0x690c3395u, 0xec9e99adu, 0x585ff075u, 0x090689d0u}};

G1Affine g1_generator() {
  Fq x = Fq::zero(), y = Fq::zero();
  x.v[0] = 1;
  y.v[0] = 2;
  return {to_mont(x), to_mont(y)};
}
G2Affine g2_generator() {
  Fq c[4];
  for (int i = 0; i < 4; i++) {
    for (int j = 0; j < 8; j++) c[i].v[j] = G2_GEN[i][j];
    c[i] = to_mont(c[i]);
  }
  return {{c[0], c[1]}, {c[2], c[3]}};
}

// table[w*256 + d] = d * 2^(8w) * G, d in 1..255 (entry 0 unused)
template <class F>
__global__ void k_fb_table(Affine<F> gen, Affine<F>* table) {
  const int w = threadIdx.x;
  XYZZ<F> b = XYZZ<F>::from_affine(gen);
  for (int i = 0; i < 8 * w; i++) b = dbl_cold(b);
  const Affine<F> base = to_affine(b);
  XYZZ<F> acc = XYZZ<F>::inf();
  table[w * 256] = Affine<F>::inf();
  for (int d = 1; d < 256; d++) {
    acc = add_mixed_cold(acc, base);
    table[w * 256 + d] = to_affine(acc);
  }
}

template <class F>
__global__ void __launch_bounds__(128)
k_fixed_base(const Affine<F>* __restrict__ table, const Fr* __restrict__ scalars, uint64_t n,
             Affine<F>* __restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    Fr s = from_mont(scalars[i]);
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int w = 0; w < 32; w++) {
      const uint32_t d = (s.v[w >> 2] >> ((w & 3) * 8)) & 0xffu;
      if (d) acc = add_mixed_cold(acc, table[w * 256 + d]);
    }
    out[i] = to_affine(acc);
  }
}

FB_HD void limbs_to_be32(const uint32_t* v, uint8_t* be) {
  for (int i = 0; i < 8; i++) {
    uint8_t* p = be + 32 - 4 * (i + 1);
    p[0] = v[i] >> 24; p[1] = v[i] >> 16; p[2] = v[i] >> 8; p[3] = v[i];
  }
}
__global__ void k_encode_g1(const G1Affine* __restrict__ in, uint64_t n, uint8_t* __restrict__ be) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    G1Affine p = in[i];
    uint8_t* o = be + i * 64;
    if (p.is_inf()) {
      for (int j = 0; j < 64; j++) o[j] = 0;
      o[0] = 0x40;
      continue;
    }
    Fq x = from_mont(p.x), y = from_mont(p.y);
    limbs_to_be32(x.v, o);
    limbs_to_be32(y.v, o + 32);
  }
}
__global__ void k_encode_g2(const G2Affine* __restrict__ in, uint64_t n, uint8_t* __restrict__ be) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    G2Affine p = in[i];
    uint8_t* o = be + i * 128;
    if (p.is_inf()) {
      for (int j = 0; j < 128; j++) o[j] = 0;
      o[0] = 0x40;
      continue;
    }
    Fq c[4] = {from_mont(p.x.c1), from_mont(p.x.c0), from_mont(p.y.c1), from_mont(p.y.c0)};
    for (int j = 0; j < 4; j++) limbs_to_be32(c[j].v, o + 32 * j);
  }
}

struct FixedBase {
  G1Affine* t1 = nullptr;
  G2Affine* t2 = nullptr;
  Fr* dsc = nullptr;
  void* dpts = nullptr;
  uint8_t* dbe = nullptr;
  uint64_t chunk = 0;
  cudaStream_t st = nullptr;

  int init(cudaStream_t s, uint64_t max_n) {
    st = s;
    chunk = std::min<uint64_t>(std::max<uint64_t>(max_n, 1), 1ull << 22);
    FB_CUDA(cudaMalloc(&t1, 32 * 256 * sizeof(G1Affine)));
    FB_CUDA(cudaMalloc(&t2, 32 * 256 * sizeof(G2Affine)));
    FB_CUDA(cudaMalloc(&dsc, chunk * sizeof(Fr)));
    FB_CUDA(cudaMalloc(&dpts, chunk * sizeof(G2Affine)));
    FB_CUDA(cudaMalloc(&dbe, chunk * 128));
    k_fb_table<Fq><<<1, 32, 0, st>>>(g1_generator(), t1);
    k_fb_table<Fq2><<<1, 32, 0, st>>>(g2_generator(), t2);
    FB_CUDA(cudaStreamSynchronize(st));
    return FB_OK;
  }
  void release() {
    cudaFree(t1); cudaFree(t2); cudaFree(dsc); cudaFree(dpts); cudaFree(dbe);
  }
  // scalars (host, Montgomery) -> points; encode: 0 raw affine, 1 bellman big-endian
  int run(int group, const Fr* scalars, uint64_t n, uint8_t* out, int encode) {
    const size_t psz = group == 1 ? 64 : 128;
    for (uint64_t off = 0; off < n; off += chunk) {
      const uint64_t cnt = std::min(chunk, n - off);
      FB_CUDA(cudaMemcpyAsync(dsc, scalars + off, cnt * sizeof(Fr), cudaMemcpyHostToDevice, st));
      const unsigned blocks = (unsigned)std::min<uint64_t>((cnt + 127) / 128, 148 * 16);
      if (group == 1) {
        k_fixed_base<Fq><<<blocks, 128, 0, st>>>(t1, dsc, cnt, (G1Affine*)dpts);
        if (encode) k_encode_g1<<<blocks, 128, 0, st>>>((const G1Affine*)dpts, cnt, dbe);
      } else {
        k_fixed_base<Fq2><<<blocks, 128, 0, st>>>(t2, dsc, cnt, (G2Affine*)dpts);
        if (encode) k_encode_g2<<<blocks, 128, 0, st>>>((const G2Affine*)dpts, cnt, dbe);
      }
      FB_CUDA(cudaMemcpyAsync(out + off * psz, encode ? (const void*)dbe : (const void*)dpts,
                              cnt * psz, cudaMemcpyDeviceToHost, st));
      FB_CUDA(cudaStreamSynchronize(st));
    }
    FB_CUDA(cudaGetLastError());
    return FB_OK;
  }
};

// host loops of the setup (powers of tau, query scalars) split over the host cores
template <class Fn>
static void host_parallel(uint64_t n, Fn fn) {
  const unsigned T = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(std::min(32u, std::max(1u, std::thread::hardware_concurrency())), n / 65536));
  if (T <= 1) { fn(0, n); return; }
  std::vector<std::thread> th;
  const uint64_t chunk = (n + T - 1) / T;
  for (unsigned t = 0; t < T; t++) {
    const uint64_t lo = t * chunk, hi = std::min(n, lo + chunk);
    if (lo >= hi) break;
    th.emplace_back([=] { fn(lo, hi); });
  }
  for (auto& x : th) x.join();
}
static hfr::H host_pow(hfr::H a, uint64_t e) {
  hfr::H r = hfr::one();
  while (e) {
    if (e & 1) r = hfr::mul(r, a);
    a = hfr::mul(a, a);
    e >>= 1;
  }
  return r;
}

static inline hfr::H H_of(const uint64_t x[4]) { hfr::H h; memcpy(h.v, x, 32); return h; }
static inline hfr::H H_of(const Fr& x) { hfr::H h; memcpy(h.v, x.v, 32); return h; }
static inline Fr F_of(const hfr::H& h) { Fr r; memcpy(r.v, h.v, 32); return r; }

}  // namespace fb

using namespace fb;

extern "C" {

int fb_test_fixed_base(fb_ctx* ctx_, int group, const uint64_t* scalars, uint64_t n, uint8_t* out_raw) try {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  if (!ctx || !scalars || !out_raw || (group != 1 && group != 2)) return FB_ERR_ARG;
  FB_CUDA(cudaSetDevice(ctx->device));
  FixedBase fbk;
  int rc = fbk.init(ctx->stream, n);
  if (!rc) rc = fbk.run(group, reinterpret_cast<const Fr*>(scalars), n, out_raw, 0);
  fbk.release();
  return rc;
} FB_ABI_CATCH_INT

// shard / nshards: nshards == 1 generates every point.  Otherwise only the points fb_pk_load_shard(shard, nshards)
// on this context keeps are generated (contiguous slices of l, a, b_g1, b_g2; for h the bit-reversed positions of
// the shard's range); every other query point is written as the point at infinity, so the output is still a
// well-formed bellman Parameters byte string of the full size -- the fixed-base work, which is all of the
// GPU time of a setup, drops by nshards.
static int setup_impl(fb_ctx* ctx_, const fb_circuit* circuit, const uint64_t trapdoor[5][4], int shard, int nshards,
                      uint8_t** params_out, size_t* len_out) {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  const Circuit* c = reinterpret_cast<const Circuit*>(circuit);
  if (!ctx || !c || !trapdoor || !params_out || !len_out || nshards < 1 || shard < 0 || shard >= nshards) {
    set_error("fb_setup: bad argument");
    return FB_ERR_ARG;
  }
  FB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const HostCsr& csr = c->csr;
  const uint32_t n_in = c->n_in, n_aux = c->n_aux, ng = csr.n_gates;
  const uint64_t n_rows = (uint64_t)ng + n_in;
  uint64_t m = 1;
  int k = 0;
  while (m < n_rows) {
    m *= 2;
    k++;
    if (k >= 28) { set_error("PolynomialDegreeTooLarge"); return FB_ERR_DOMAIN; }
  }
  if (k == 0) { m = 2; k = 1; }
  const hfr::H alpha = H_of(trapdoor[0]), beta = H_of(trapdoor[1]), gamma = H_of(trapdoor[2]),
               delta = H_of(trapdoor[3]), tau = H_of(trapdoor[4]);
  if (hfr::is_zero(gamma) || hfr::is_zero(delta)) { set_error("gamma/delta must be non-zero"); return FB_ERR_ARG; }
  // powers of tau, h scalars
  std::vector<hfr::H> pw(m);
  host_parallel(m, [&](uint64_t lo, uint64_t hi) {
    hfr::H u = host_pow(tau, lo);
    for (uint64_t i = lo; i < hi; i++) { pw[i] = u; u = hfr::mul(u, tau); }
  });
  const hfr::H z_tau = hfr::sub(hfr::mul(pw[m - 1], tau), hfr::one());
  const hfr::H dinv = hfr::inv(delta), ginv = hfr::inv(gamma);
  const hfr::H hcoef = hfr::mul(z_tau, dinv);
  std::vector<Fr> h_s(m - 1);
  host_parallel(m - 1, [&](uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; i++) h_s[i] = F_of(hfr::mul(pw[i], hcoef));
  });
  // Lagrange values at tau: ifft of the powers (on the GPU)
  {
    NttDomain dom;
    if (dom.init(k, st) != 0) { set_error("domain init failed"); return FB_ERR_CUDA; }
    struct Dev {  // freed on every exit path
      Fr* p = nullptr;
      ~Dev() { if (p) cudaFree(p); }
    } dx, ds;
    struct DomGuard {
      NttDomain& d;
      ~DomGuard() { d.destroy(); }
    } dom_guard{dom};
    FB_CUDA(cudaMalloc(&dx.p, m * sizeof(Fr)));
    FB_CUDA(cudaMalloc(&ds.p, m * sizeof(Fr)));
    FB_CUDA(cudaMemcpyAsync(dx.p, pw.data(), m * sizeof(Fr), cudaMemcpyHostToDevice, st));
    dom.transform(dx.p, ds.p, 1, st);
    FB_CUDA(cudaMemcpyAsync(pw.data(), dx.p, m * sizeof(Fr), cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaStreamSynchronize(st));
  }
  const std::vector<hfr::H>& lag = pw;
  // column accumulation: at[k] = sum_rows A[row,k] L_row(tau), ...
  const uint32_t nv = n_in + n_aux;
  std::vector<hfr::H> acc[3];
  {  // the three matrices accumulate into their own arrays: one host thread each
    std::vector<std::thread> th;
    for (int mi = 0; mi < 3; mi++)
      th.emplace_back([&, mi] {
        acc[mi].assign(nv, hfr::zero());
        for (uint32_t row = 0; row < ng; row++) {
          const hfr::H& lj = lag[row];
          for (uint32_t p = csr.rowptr[mi][row]; p < csr.rowptr[mi][row + 1]; p++) {
            const uint32_t ci = csr.cidx[mi][p];
            hfr::H& dst = acc[mi][csr.col[mi][p]];
            if (ci == 0) dst = hfr::add(dst, lj);
            else if (ci == 1) dst = hfr::sub(dst, lj);
            else dst = hfr::add(dst, hfr::mul(H_of(csr.coef[ci - 2]), lj));
          }
        }
      });
    for (auto& x : th) x.join();
  }
  for (uint32_t i = 0; i < n_in; i++) acc[0][i] = hfr::add(acc[0][i], lag[ng + i]);
  std::vector<Fr> ic_s(n_in), l_s(n_aux), a_s, b_s;
  host_parallel(nv, [&](uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; i++) {
      hfr::H t = hfr::add(hfr::add(hfr::mul(beta, acc[0][i]), hfr::mul(alpha, acc[1][i])), acc[2][i]);
      if (i < n_in) ic_s[i] = F_of(hfr::mul(t, ginv));
      else l_s[i - n_in] = F_of(hfr::mul(t, dinv));
    }
  });
  for (uint32_t i = 0; i < nv; i++) {  // points at infinity are filtered out of the a / b queries
    if (!hfr::is_zero(acc[0][i])) a_s.push_back(F_of(acc[0][i]));
    if (!hfr::is_zero(acc[1][i])) b_s.push_back(F_of(acc[1][i]));
  }
  for (int mi = 0; mi < 3; mi++) std::vector<hfr::H>().swap(acc[mi]);
  std::vector<hfr::H>().swap(pw);
  // layout of the output
  const uint64_t n_h = m - 1, n_a = a_s.size(), n_b = b_s.size();
  const size_t total = 64 + 64 + 128 + 128 + 64 + 128 + 4 + (size_t)n_in * 64 + 4 + n_h * 64 + 4 +
                       (size_t)n_aux * 64 + 4 + n_a * 64 + 4 + n_b * 64 + 4 + n_b * 128;
  uint8_t* out = (uint8_t*)malloc(total);
  if (!out) { set_error("out of host memory for %zu bytes of Parameters", total); return FB_ERR_ARG; }
  FixedBase fbk;
  int rc = fbk.init(st, std::max<uint64_t>(std::max<uint64_t>(n_h, n_aux), std::max(n_a, n_b)));
  size_t pos = 0;
  auto len_be = [&](uint64_t n) {
    out[pos] = (uint8_t)(n >> 24); out[pos + 1] = (uint8_t)(n >> 16);
    out[pos + 2] = (uint8_t)(n >> 8); out[pos + 3] = (uint8_t)n;
    pos += 4;
  };
  auto emit = [&](int group, const Fr* s, uint64_t n) {
    if (!rc) rc = fbk.run(group, s, n, out + pos, 1);
    pos += n * (group == 1 ? 64 : 128);
  };
  auto fill_inf = [&](uint8_t* p, uint64_t n, size_t psz) {
    memset(p, 0, n * psz);
    for (uint64_t i = 0; i < n; i++) p[i * psz] = 0x40;
  };
  // a query of which this shard keeps the slice [n*shard/nshards, n*(shard+1)/nshards)
  auto emit_slice = [&](int group, const Fr* s, uint64_t n) {
    const size_t psz = group == 1 ? 64 : 128;
    if (nshards == 1) { emit(group, s, n); return; }
    const uint64_t lo = n * shard / nshards, hi = n * (shard + 1) / nshards;
    fill_inf(out + pos, lo, psz);
    if (!rc) rc = fbk.run(group, s + lo, hi - lo, out + pos + lo * psz, 1);
    fill_inf(out + pos + hi * psz, n - hi, psz);
    pos += n * psz;
  };
  // h: the loader keeps positions [h_lo, h_lo + h_cnt) of the BIT-REVERSED array, i.e. coefficient i = brev(p)
  auto emit_h = [&](const Fr* s, uint64_t n) {
    if (nshards == 1) { emit(1, s, n); return; }
    uint64_t h_lo, h_cnt;
    key_h_range(m, key_dist_g(ctx, k, shard, nshards), shard, nshards, &h_lo, &h_cnt);
    std::vector<Fr> sel(h_cnt);
    std::vector<uint32_t> idx(h_cnt);
    for (uint64_t p = 0; p < h_cnt; p++) {
      uint64_t i = 0, q = h_lo + p;
      for (int b = 0; b < k; b++) i |= ((q >> b) & 1) << (k - 1 - b);
      idx[p] = (uint32_t)i;      // i < m - 1 for every position the loader reads (position of m-1 is m-1 itself)
      sel[p] = i < n ? s[i] : Fr::zero();
    }
    std::vector<uint8_t> pts(h_cnt * 64);
    if (!rc) rc = fbk.run(1, sel.data(), h_cnt, pts.data(), 1);
    fill_inf(out + pos, n, 64);
    for (uint64_t p = 0; p < h_cnt; p++)
      if (idx[p] < n) memcpy(out + pos + (uint64_t)idx[p] * 64, pts.data() + p * 64, 64);
    pos += n * 64;
  };
  Fr vk1[3] = {F_of(alpha), F_of(beta), F_of(delta)};
  Fr vk2[3] = {F_of(beta), F_of(gamma), F_of(delta)};
  emit(1, &vk1[0], 1);  // alpha_g1
  emit(1, &vk1[1], 1);  // beta_g1
  emit(2, &vk2[0], 1);  // beta_g2
  emit(2, &vk2[1], 1);  // gamma_g2
  emit(1, &vk1[2], 1);  // delta_g1
  emit(2, &vk2[2], 1);  // delta_g2
  len_be(n_in); emit(1, ic_s.data(), n_in);
  len_be(n_h); emit_h(h_s.data(), n_h);
  len_be(n_aux); emit_slice(1, l_s.data(), n_aux);
  len_be(n_a); emit_slice(1, a_s.data(), n_a);
  len_be(n_b); emit_slice(1, b_s.data(), n_b);
  len_be(n_b); emit_slice(2, b_s.data(), n_b);
  fbk.release();
  if (rc) { free(out); return rc; }
  *params_out = out;
  *len_out = total;
  return FB_OK;
}

int fb_setup(fb_ctx* ctx, const fb_circuit* circuit, const uint64_t trapdoor[5][4], uint8_t** params_out,
             size_t* len_out) try {
  return setup_impl(ctx, circuit, trapdoor, 0, 1, params_out, len_out);
} FB_ABI_CATCH_INT

int fb_setup_shard(fb_ctx* ctx, const fb_circuit* circuit, const uint64_t trapdoor[5][4], int shard, int nshards,
                   uint8_t** params_out, size_t* len_out) try {
  return setup_impl(ctx, circuit, trapdoor, shard, nshards, params_out, len_out);
} FB_ABI_CATCH_INT

}  // extern "C"
