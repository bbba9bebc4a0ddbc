// C ABI of libfawkes_b200.so -- see include/fawkes_b200.h for the reference interfaces
// each entry point replaces.
#include "../../include/fawkes_b200.h"

#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <future>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <thread>

#include "internal.h"

namespace fb {
const char* last_error_cstr();
void kstat_enable(bool on);
void kstat_reset();
void kstat_collect(int kind, unsigned long long* launches, double* ms);

static uint32_t be32(const uint8_t* p) {
  return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

int parse_params(const uint8_t* b, size_t len, ParamsView& v) {
  size_t pos = 0;
  auto take = [&](size_t n) -> const uint8_t* {
    if (pos + n > len) return nullptr;
    const uint8_t* p = b + pos;
    pos += n;
    return p;
  };
  if (!(v.alpha_g1 = take(64)) || !(v.beta_g1 = take(64)) || !(v.beta_g2 = take(128)) ||
      !(v.gamma_g2 = take(128)) || !(v.delta_g1 = take(64)) || !(v.delta_g2 = take(128))) {
    set_error("Parameters truncated in verifying key");
    return FB_ERR_FORMAT;
  }
  struct { const uint8_t** p; uint32_t* n; size_t sz; } secs[6] = {
      {&v.ic, &v.n_ic, 64}, {&v.h, &v.n_h, 64}, {&v.l, &v.n_l, 64},
      {&v.a, &v.n_a, 64},   {&v.b1, &v.n_b1, 64}, {&v.b2, &v.n_b2, 128}};
  for (auto& s : secs) {
    const uint8_t* q = take(4);
    if (!q) { set_error("Parameters truncated at a length field"); return FB_ERR_FORMAT; }
    *s.n = be32(q);
    if (!(*s.p = take((size_t)*s.n * s.sz))) {
      set_error("Parameters truncated inside a query (%u points announced)", *s.n);
      return FB_ERR_FORMAT;
    }
  }
  return FB_OK;
}

__global__ void k_gather_bitrev(G1Affine* __restrict__ dst, const G1Affine* __restrict__ src, int k,
                                uint32_t lo, uint32_t cnt) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
    uint32_t pos = lo + i;
    uint32_t r = __brev(pos) >> (32 - k);
    dst[i] = src[r];
  }
}

static void free_host_tables(void* p);
static void free_batch(void* p);
static void* build_host_tables(const ProvingKey* pk);
static void pk_release(ProvingKey* pk) {
  if (!pk) return;
  for (ProvingKey* s : pk->slots) pk_release(s);
  pk->slots.clear();
  if (!pk->is_slot) {  // a slot only borrows the immutable arrays
    cudaFree(pk->h); cudaFree(pk->l); cudaFree(pk->a); cudaFree(pk->b1); cudaFree(pk->b2);
    cudaFree(pk->a_map); cudaFree(pk->b_map);
    free_csr(pk->csr);
    pk->dom.destroy();
  }
  cudaFree(pk->w);
  for (int i = 0; i < 3; i++) { cudaFree(pk->ev[i]); cudaFree(pk->xtmp[i]); }
  cudaFree(pk->scratch);
  for (auto& x : pk->msm) x.release();
  cudaFree(pk->results);
  if (pk->results_host) cudaFreeHost(pk->results_host);
  for (auto& e : pk->msm_done) if (e) cudaEventDestroy(e);
  if (pk->batch) free_batch(pk->batch);
  if (pk->graph_exec) cudaGraphExecDestroy(pk->graph_exec);
  if (pk->w_stage) cudaFreeHost(pk->w_stage);
  if (pk->is_slot) {
    Ctx* c = pk->ctx;
    if (c) {
      cudaStreamDestroy(c->stream);
      for (int i = 0; i < 3; i++) { cudaStreamDestroy(c->aux[i]); cudaEventDestroy(c->aux_done[i]); }
      if (c->stage) {
        cudaFreeHost(c->stage);
        for (auto& s : c->copy_stream) cudaStreamDestroy(s);
        for (auto& e : c->stage_ev) cudaEventDestroy(e);
        for (auto& e : c->copy_done) cudaEventDestroy(e);
      }
      delete c;
    }
  } else if (pk->host_tables) {
    free_host_tables(pk->host_tables);
  }
  delete pk;
}

struct Timing {
  cudaEvent_t ev[6];
  float ms[6] = {0, 0, 0, 0, 0, 0};
  bool init = false;
  ~Timing() {  // thread_local: the batch worker threads release their events on exit
    if (init) for (auto& e : ev) cudaEventDestroy(e);
  }
};
static thread_local Timing g_timing;  // per calling thread (fb_prove_batch proves on several)

// log2 of the number of ranks the R1CS rows and the H pipeline of a key are sharded over (0 = replicated):
// world = nshards = 2^g ranks with an NCCL exchange, shard = rank, and a domain large enough.
int key_dist_g(const Ctx* ctx, int k, int shard, int nshards) {
  if (!(nshards > 1 && ctx->exchange && ctx->world == nshards && ctx->rank == shard && (nshards & (nshards - 1)) == 0))
    return 0;
  int g = 0;
  while ((1 << g) < nshards) g++;
  NttDomain probe;
  probe.k = k;
  const char* env = getenv("FB_DIST_NTT_MIN_LOG");
  const int min_log = env ? atoi(env) : 20;
  return (probe.dist_supported(g) && k >= min_log) ? g : 0;
}
// positions [lo, lo + cnt) of the bit-reversed h array a shard keeps
void key_h_range(uint64_t m, int dist_g, int shard, int nshards, uint64_t* lo, uint64_t* cnt) {
  const uint64_t nh = m - 1;
  if (dist_g) {  // H comes out of the distributed pipeline in block layout: positions [rank*ml, (rank+1)*ml)
    const uint64_t ml = m >> dist_g;
    *lo = (uint64_t)shard * ml;
    *cnt = std::min<uint64_t>(ml, nh - *lo);
  } else {
    *lo = nh * shard / nshards;
    *cnt = nh * (shard + 1) / nshards - *lo;
  }
}

static int load_key_once(Ctx* ctx, const uint8_t* params, size_t len, const Circuit* circ, int checked,
                         int shard, int nshards, bool allow_tables, bool* used_tables, ProvingKey** out) {
  if (!ctx || !params || !circ || !out || nshards < 1 || shard < 0 || shard >= nshards) {
    set_error("fb_pk_load: bad argument");
    return FB_ERR_ARG;
  }
  FB_CUDA(cudaSetDevice(ctx->device));
  ParamsView v;
  int rc = parse_params(params, len, v);
  if (rc) return rc;
  if (v.n_ic != circ->n_in || v.n_l != circ->n_aux) {
    set_error("Parameters are for n_in=%u n_aux=%u but the circuit has n_in=%u n_aux=%u", v.n_ic,
              v.n_l, circ->n_in, circ->n_aux);
    return FB_ERR_ARG;
  }
  if (v.n_b1 != v.n_b2) { set_error("b_g1 and b_g2 lengths differ"); return FB_ERR_FORMAT; }
  ProvingKey* pk = new ProvingKey();
  pk->ctx = ctx;
  pk->n_in = circ->n_in;
  pk->n_aux = circ->n_aux;
  pk->shard = shard;
  pk->nshards = nshards;
  const HostCsr& csr = circ->csr;
  pk->n_rows = csr.n_gates + pk->n_in;
  // domain (bellman EvaluationDomain::from_coeffs)
  uint64_t m = 1;
  int k = 0;
  while (m < pk->n_rows) {
    m *= 2;
    k++;
    if (k >= 28) { delete pk; set_error("PolynomialDegreeTooLarge"); return FB_ERR_DOMAIN; }
  }
  if (k == 0) { m = 2; k = 1; }  // degenerate single-row circuit: pad to 2
  pk->k = k;
  pk->m = m;
  if (v.n_h + 1 < m) {
    delete pk;
    set_error("h query has %u points, need %llu", v.n_h, (unsigned long long)(m - 1));
    return FB_ERR_FORMAT;
  }
  // distributed H pipeline: world = nshards = 2^g ranks with an NCCL exchange and a domain large enough
  const int dist_g = key_dist_g(ctx, k, shard, nshards);
  pk->dist_g = dist_g;
  const uint64_t ml = dist_g ? (m >> dist_g) : m;  // local length of the evaluation arrays
  cudaStream_t st = ctx->stream;
#define PK_TRY(x) do { int _r = (x); if (_r) { pk_release(pk); return _r; } } while (0)
#define PK_CUDA(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { set_error("%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(_e)); pk_release(pk); return FB_ERR_CUDA; } } while (0)
  // vk points on host: always checked, as bellman's VerifyingKey::read does; ic points may not be at infinity
  {
    struct { const char* name; int rc; } vkp[5] = {
        {"alpha_g1", host_decode_g1(v.alpha_g1, pk->alpha_g1)}, {"beta_g1", host_decode_g1(v.beta_g1, pk->beta_g1)},
        {"delta_g1", host_decode_g1(v.delta_g1, pk->delta_g1)}, {"beta_g2", host_decode_g2(v.beta_g2, pk->beta_g2)},
        {"delta_g2", host_decode_g2(v.delta_g2, pk->delta_g2)}};
    G2Affine gamma;
    int grc = host_decode_g2(v.gamma_g2, gamma);
    for (auto& e : vkp)
      if (e.rc) { pk_release(pk); set_error("invalid verifying-key point %s (%s)", e.name, point_error(e.rc)); return FB_ERR_FORMAT; }
    if (grc) { pk_release(pk); set_error("invalid verifying-key point gamma_g2 (%s)", point_error(grc)); return FB_ERR_FORMAT; }
    for (uint32_t i = 0; i < v.n_ic; i++) {
      G1Affine icp;
      const int irc = host_decode_g1(v.ic + 64 * (size_t)i, icp, FB_LOAD_NO_INFINITY);
      if (irc) { pk_release(pk); set_error("invalid verifying-key point ic[%u] (%s)", i, point_error(irc)); return FB_ERR_FORMAT; }
    }
  }
  // structural density (bellman DensityTracker: a variable counts when it appears)
  std::vector<uint8_t> a_d(pk->n_in + pk->n_aux, 0), b_d(pk->n_in + pk->n_aux, 0);
  for (uint32_t c : csr.col[0]) a_d[c] = 1;
  for (uint32_t c : csr.col[1]) b_d[c] = 1;
  std::vector<uint32_t> a_map, b_map;
  for (uint32_t i = 0; i < pk->n_in; i++) a_map.push_back(i);  // inputs: full density
  for (uint32_t i = 0; i < pk->n_aux; i++) if (a_d[pk->n_in + i]) a_map.push_back(pk->n_in + i);
  for (uint32_t i = 0; i < pk->n_in + pk->n_aux; i++) if (b_d[i]) b_map.push_back(i);
  if (a_map.size() != v.n_a || b_map.size() != v.n_b1) {
    set_error("query/density mismatch: a has %u points for %zu dense variables, b has %u for %zu",
              v.n_a, a_map.size(), v.n_b1, b_map.size());
    pk_release(pk);
    return FB_ERR_DENSITY;
  }
  auto slice = [&](uint64_t n, uint64_t& lo, uint64_t& cnt) {
    lo = n * shard / nshards;
    cnt = n * (shard + 1) / nshards - lo;
  };
  uint64_t lo, cnt;
  // MSM plans first: in window-table mode every base array is allocated W times larger
  {
    uint64_t cnt_h, cnt_l, cnt_a, cnt_b;
    slice(m - 1, lo, cnt_h);
    if (dist_g) cnt_h = std::min<uint64_t>(ml, (m - 1) - (uint64_t)shard * ml);
    slice(v.n_l, lo, cnt_l);
    slice(v.n_a, lo, cnt_a);
    slice(v.n_b1, lo, cnt_b);
    bool tables = g_msm_tables == 1;
    if (g_msm_tables < 0 && allow_tables) {
      const MsmPlan th = MsmPlan::make((uint32_t)cnt_h, true), tl = MsmPlan::make((uint32_t)cnt_l, true),
                    ta = MsmPlan::make((uint32_t)cnt_a, true), tb = MsmPlan::make((uint32_t)cnt_b, true);
      const uint64_t need = (th.table_points() + tl.table_points() + ta.table_points()) * 64 + tb.table_points() * 192;
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      // everything else a key holds (CSR, twiddles, workspaces, MSM scratch) is ~25 x 32 B per row
      tables = need + (uint64_t)m * 32 * 25 + (8ull << 30) < free_b;
    }
    *used_tables = tables;
    pk->plan_h = MsmPlan::make((uint32_t)cnt_h, tables);
    pk->plan_l = MsmPlan::make((uint32_t)cnt_l, tables);
    pk->plan_a = MsmPlan::make((uint32_t)cnt_a, tables);
    pk->plan_b = MsmPlan::make((uint32_t)cnt_b, tables);
    pk->table_bytes = tables ? (pk->plan_h.table_points() - cnt_h + pk->plan_l.table_points() - cnt_l +
                                pk->plan_a.table_points() - cnt_a) * 64 + (pk->plan_b.table_points() - cnt_b) * 192
                             : 0;
  }
  // h: decode all to a temporary, gather this shard's bit-reversed positions
  {
    const uint64_t nh = m - 1;
    G1Affine* tmp = nullptr;
    PK_CUDA(cudaMalloc(&tmp, m * sizeof(G1Affine)));
    PK_CUDA(cudaMemsetAsync(tmp, 0, m * sizeof(G1Affine), st));
    rc = decode_g1_be(v.h, nh, tmp, checked, st);
    if (rc) { cudaFree(tmp); pk_release(pk); return rc; }
    slice(nh, lo, cnt);
    if (dist_g) {  // H comes out of the distributed pipeline in block layout: positions [rank*ml, (rank+1)*ml)
      lo = (uint64_t)shard * ml;
      cnt = std::min<uint64_t>(ml, nh - lo);
    }
    pk->len_h = (uint32_t)cnt;
    PK_CUDA(cudaMalloc(&pk->h, std::max<uint64_t>(pk->plan_h.table_points(), 1) * sizeof(G1Affine)));
    if (cnt) k_gather_bitrev<<<(unsigned)std::min<uint64_t>((cnt + 255) / 256, 148 * 16), 256, 0, st>>>(
        pk->h, tmp, k, (uint32_t)lo, (uint32_t)cnt);
    PK_CUDA(cudaStreamSynchronize(st));
    cudaFree(tmp);
  }
  slice(v.n_l, lo, cnt);
  PK_CUDA(cudaMalloc(&pk->l, std::max<uint64_t>(pk->plan_l.table_points(), 1) * sizeof(G1Affine)));
  PK_TRY(decode_g1_be(v.l + lo * 64, cnt, pk->l, checked, st));
  slice(v.n_a, lo, cnt);
  pk->len_a = (uint32_t)cnt;
  PK_CUDA(cudaMalloc(&pk->a, std::max<uint64_t>(pk->plan_a.table_points(), 1) * sizeof(G1Affine)));
  PK_TRY(decode_g1_be(v.a + lo * 64, cnt, pk->a, checked, st));
  PK_CUDA(cudaMalloc(&pk->a_map, std::max<uint64_t>(cnt, 1) * 4));
  PK_CUDA(cudaMemcpyAsync(pk->a_map, a_map.data() + lo, cnt * 4, cudaMemcpyHostToDevice, st));
  PK_CUDA(cudaStreamSynchronize(st));
  slice(v.n_b1, lo, cnt);
  pk->len_b = (uint32_t)cnt;
  PK_CUDA(cudaMalloc(&pk->b1, std::max<uint64_t>(pk->plan_b.table_points(), 1) * sizeof(G1Affine)));
  PK_TRY(decode_g1_be(v.b1 + lo * 64, cnt, pk->b1, checked, st));
  PK_CUDA(cudaMalloc(&pk->b2, std::max<uint64_t>(pk->plan_b.table_points(), 1) * sizeof(G2Affine)));
  PK_TRY(decode_g2_be(v.b2 + lo * 128, cnt, pk->b2, checked, st));
  PK_CUDA(cudaMalloc(&pk->b_map, std::max<uint64_t>(cnt, 1) * 4));
  PK_CUDA(cudaMemcpyAsync(pk->b_map, b_map.data() + lo, cnt * 4, cudaMemcpyHostToDevice, st));
  PK_CUDA(cudaStreamSynchronize(st));
  // CSR + domain + workspaces
  if (dist_g) {
    HostCsr sub;
    const uint32_t G = 1u << dist_g;
    for (int mi = 0; mi < 3; mi++) sub.rowptr[mi].push_back(0);
    for (uint32_t row = shard; row < csr.n_gates; row += G) {
      for (int mi = 0; mi < 3; mi++) {
        for (uint32_t q = csr.rowptr[mi][row]; q < csr.rowptr[mi][row + 1]; q++) {
          sub.col[mi].push_back(csr.col[mi][q]);
          sub.cidx[mi].push_back(csr.cidx[mi][q]);
        }
        sub.rowptr[mi].push_back((uint32_t)sub.col[mi].size());
      }
      sub.n_gates++;
    }
    sub.coef = csr.coef;
    PK_TRY(upload_csr(sub, pk->csr, st));
    pk->n_gates_global = csr.n_gates;
  } else {
    PK_TRY(upload_csr(csr, pk->csr, st));
    pk->n_gates_global = csr.n_gates;
  }
  if (pk->dom.init(k, st) != 0) {
    set_error("NTT domain init failed: %s", cudaGetErrorString(cudaGetLastError()));
    pk_release(pk);
    return FB_ERR_CUDA;
  }
  // padded to a multiple of the world size so the witness can be all-gathered in equal chunks
  PK_CUDA(cudaMalloc(&pk->w, ((size_t)(pk->n_in + pk->n_aux) + (size_t)nshards) * sizeof(Fr)));
  for (int i = 0; i < 3; i++) PK_CUDA(cudaMalloc(&pk->ev[i], ml * sizeof(Fr)));
  PK_CUDA(cudaMalloc(&pk->scratch, ml * sizeof(Fr)));
  if (dist_g)
    for (int i = 0; i < 3; i++) PK_CUDA(cudaMalloc(&pk->xtmp[i], ml * sizeof(Fr)));
  if (pk->plan_h.n != pk->len_h || pk->plan_a.n != pk->len_a || pk->plan_b.n != pk->len_b) {
    set_error("internal: MSM plan sizes disagree with the loaded shard");
    pk_release(pk);
    return FB_ERR_ARG;
  }
  // window tables: 2^(c w) P for every window, next to the bases (once per key)
  PK_TRY(msm_build_table_g1(pk->h, pk->plan_h, st) ? FB_ERR_CUDA : 0);
  PK_TRY(msm_build_table_g1(pk->l, pk->plan_l, st) ? FB_ERR_CUDA : 0);
  PK_TRY(msm_build_table_g1(pk->a, pk->plan_a, st) ? FB_ERR_CUDA : 0);
  PK_TRY(msm_build_table_g1(pk->b1, pk->plan_b, st) ? FB_ERR_CUDA : 0);
  PK_TRY(msm_build_table_g2(pk->b2, pk->plan_b, st) ? FB_ERR_CUDA : 0);
  PK_CUDA(cudaStreamSynchronize(st));
  const MsmPlan msm_plans[4] = {pk->plan_h, pk->plan_l, pk->plan_a, pk->plan_b};
  int arc = 0;
  for (int i = 0; i < 4 && !arc; i++) arc = pk->msm[i].alloc(&msm_plans[i], 1, i == 3);
  if (arc != 0) {
    set_error("MSM scratch allocation failed");
    pk_release(pk);
    return FB_ERR_CUDA;
  }
  PK_CUDA(cudaMalloc(&pk->results, 5 * MSM_VBITS * sizeof(G2XYZZ)));
  for (auto& e : pk->msm_done) PK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  PK_CUDA(cudaMallocHost(&pk->results_host, 5 * MSM_VBITS * sizeof(G2XYZZ)));
  PK_CUDA(cudaStreamSynchronize(st));
  pk->host_tables = build_host_tables(pk);
  *out = pk;
  return FB_OK;
#undef PK_TRY
#undef PK_CUDA
}

// In auto mode the window tables are a best-effort use of free HBM: if the estimate was too optimistic and
// an allocation fails, the key is loaded again with plain 64-byte bases instead of failing the load.
static int load_key(Ctx* ctx, const uint8_t* params, size_t len, const Circuit* circ, int checked,
                    int shard, int nshards, ProvingKey** out) {
  bool used_tables = false;
  int rc = load_key_once(ctx, params, len, circ, checked, shard, nshards, true, &used_tables, out);
  if (rc == FB_ERR_CUDA && used_tables && g_msm_tables < 0) {
    cudaGetLastError();
    rc = load_key_once(ctx, params, len, circ, checked, shard, nshards, false, &used_tables, out);
  }
  return rc;
}

// ------------------------------------------------------------- assembly ---
static void fr_canonical(const uint64_t x[4], uint32_t out[8]) {
  Fr t;
  memcpy(t.v, x, 32);
  t = from_mont(t);
  memcpy(out, t.v, 32);
}

// SURVEY.md App. C.5:  A = alpha + r*delta + sum_a ;  B = beta2 + s*delta2 + sum_b2 ;
// C = rs*delta + s*alpha + r*beta1 + s*sum_a + r*sum_b1 + sum_h + sum_l
struct VkPoints {
  G1Affine alpha_g1, beta_g1, delta_g1;
  G2Affine beta_g2, delta_g2;
};
typedef XYZZ<HFq> H1;   // host G1 accumulator (64-bit limbs)
typedef XYZZ<HFq2> H2;
static H1 h1_from(const G1Affine& p) { return p.is_inf() ? H1::inf() : H1{HFq::from(p.x), HFq::from(p.y), HFq::one(), HFq::one()}; }
static H2 h2_from(const G2Affine& p) { return p.is_inf() ? H2::inf() : H2{HFq2::from(p.x), HFq2::from(p.y), HFq2::one(), HFq2::one()}; }
static G1Affine h1_affine(const H1& p) { Affine<HFq> a = to_affine(p); return {a.x.to(), a.y.to()}; }
static G2Affine h2_affine(const H2& p) { Affine<HFq2> a = to_affine(p); return {a.x.to(), a.y.to()}; }

// Fixed-base table of one key point for the host: T[w][d-1] = d * 16^w * P, so k * P is 64 additions
// and no doublings.  Built once per key for delta_1, alpha_1, beta_1 (G1) and delta_2 (G2).
template <class HP>
struct HostTable {
  std::vector<HP> t;  // 64 * 15
  void build(const HP& p) {
    t.resize(64 * 15);
    HP base = p;
    for (int w = 0; w < 64; w++) {
      HP acc = base;
      for (int d = 1; d <= 15; d++) {
        t[w * 15 + d - 1] = acc;
        acc = add(acc, base);
      }
      base = acc;  // 16 * base
    }
  }
  HP mul(const uint32_t* k) const {  // k: 8 canonical 32-bit limbs
    HP acc = HP::inf();
    for (int w = 0; w < 64; w++) {
      const uint32_t d = (k[w >> 3] >> ((w & 7) * 4)) & 15u;
      if (d) acc = add(acc, t[w * 15 + d - 1]);
    }
    return acc;
  }
};
struct KeyTables {
  HostTable<H1> delta1, alpha1, beta1;
  HostTable<H2> delta2;
};

struct FixedTerms {  // the parts of A, B, C that depend on r, s and the key only
  H1 a;   // alpha + r*delta
  H2 b;   // beta2 + s*delta2
  H1 c;   // rs*delta + s*alpha + r*beta1
  uint32_t rc[8], sc[8];
};
static int fixed_terms(const VkPoints* pk, const KeyTables* kt, const uint64_t r[4], const uint64_t s[4],
                       FixedTerms& f) {
  if (pk->delta_g1.is_inf() || pk->delta_g2.is_inf()) {
    set_error("UnexpectedIdentity: delta is the point at infinity");
    return FB_ERR_IDENTITY;
  }
  uint32_t rsc[8];
  fr_canonical(r, f.rc);
  fr_canonical(s, f.sc);
  Fr rm, sm;
  memcpy(rm.v, r, 32);
  memcpy(sm.v, s, 32);
  Fr rs = from_mont(mul(rm, sm));
  memcpy(rsc, rs.v, 32);
  H1 al = h1_from(pk->alpha_g1);
  if (kt) {
    f.a = add(kt->delta1.mul(f.rc), al);
    f.b = add(kt->delta2.mul(f.sc), h2_from(pk->beta_g2));
    f.c = kt->delta1.mul(rsc);
    f.c = add(f.c, kt->alpha1.mul(f.sc));
    f.c = add(f.c, kt->beta1.mul(f.rc));
    return FB_OK;
  }
  H1 d1 = h1_from(pk->delta_g1), be1 = h1_from(pk->beta_g1);
  H2 d2 = h2_from(pk->delta_g2);
  f.a = add(scalar_mul(d1, f.rc), al);
  f.b = add(scalar_mul(d2, f.sc), h2_from(pk->beta_g2));
  f.c = scalar_mul(d1, rsc);
  f.c = add(f.c, scalar_mul(al, f.sc));
  f.c = add(f.c, scalar_mul(be1, f.rc));
  return FB_OK;
}
static void finish_proof(const FixedTerms& f, const H1& H, const H1& L, const H1& A, const H1& B1,
                         const H2& B2, const H1* sA, const H1* rB1, uint8_t proof_raw[256]) {
  H1 g_a = add(f.a, A);
  H2 g_b = add(f.b, B2);
  H1 g_c = add(f.c, sA ? *sA : scalar_mul(A, f.sc));
  g_c = add(g_c, rB1 ? *rB1 : scalar_mul(B1, f.rc));
  g_c = add(g_c, H);
  g_c = add(g_c, L);
  G1Affine pa = h1_affine(g_a), pc = h1_affine(g_c);
  G2Affine pb = h2_affine(g_b);
  memcpy(proof_raw, &pa, 64);
  memcpy(proof_raw + 64, &pb, 128);
  memcpy(proof_raw + 192, &pc, 64);
}
static int assemble(const VkPoints* pk, const H1& H, const H1& L, const H1& A, const H1& B1, const H2& B2,
                    const uint64_t r[4], const uint64_t s[4], uint8_t proof_raw[256]) {
  FixedTerms f;
  int rc = fixed_terms(pk, nullptr, r, s, f);
  if (rc) return rc;
  finish_proof(f, H, L, A, B1, B2, nullptr, nullptr, proof_raw);
  return FB_OK;
}

static void free_host_tables(void* p) { delete reinterpret_cast<KeyTables*>(p); }
static void* build_host_tables(const ProvingKey* pk) {
  if (pk->delta_g1.is_inf() || pk->delta_g2.is_inf()) return nullptr;  // prove reports UnexpectedIdentity
  KeyTables* kt = new KeyTables();
  kt->delta1.build(h1_from(pk->delta_g1));
  kt->alpha1.build(h1_from(pk->alpha_g1));
  kt->beta1.build(h1_from(pk->beta_g1));
  kt->delta2.build(h2_from(pk->delta_g2));
  return kt;
}

int g_msm_tables = -1;
static bool g_serial = false;  // one stream, MSMs back to back (used for per-kernel timing)
static std::atomic<int> g_prove_graph{-1}; // CUDA-graph replay for small keys: -1 = on unless FB_PROVE_GRAPH=0, 0 off, 1 on
bool kstat_enabled();

// Result slots: five arrays of MSM_VBITS bit sums (G2-sized slots), order H L A B1 B2.
static inline G2XYZZ* vslot(void* base, int i) { return reinterpret_cast<G2XYZZ*>(base) + (size_t)i * MSM_VBITS; }

// device part: w already in pk->w (on the main stream).  Every MSM leaves its bit sums in
// pk->results_host and records an event on its stream.  NOT synchronised on return.
static int prove_launch(ProvingKey* pk, uint64_t* h_out, const Fr* w = nullptr) {
  if (!w) w = pk->w;  // fb_prove_device: the caller's device buffer is read in place, no copy
  Ctx* ctx = pk->ctx;
  cudaStream_t st = ctx->stream;
  cudaStream_t sL = g_serial ? st : ctx->aux[0], sA = g_serial ? st : ctx->aux[1],
               sB = g_serial ? st : ctx->aux[2];
  Timing& T = g_timing;
  const uint64_t m = pk->m;
  cudaEventRecord(T.ev[1], st);  // witness resident
  if (!g_serial) {
    for (int i = 0; i < 3; i++) FB_CUDA(cudaStreamWaitEvent(ctx->aux[i], T.ev[1], 0));
  }
  const uint64_t nh = m - 1;
  const uint64_t h_lo = nh * pk->shard / pk->nshards;
  const uint64_t l_lo = (uint64_t)pk->n_aux * pk->shard / pk->nshards;
  auto fetch = [&](int slot, size_t bytes, cudaStream_t s, int ev) -> cudaError_t {
    cudaError_t e = cudaMemcpyAsync(vslot(pk->results_host, slot), vslot(pk->results, slot), bytes,
                                    cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaEventRecord(pk->msm_done[ev], s);
    return e;
  };
  // witness-only MSMs start at once on their own streams; the G2 one first (longest)
  int rc = msm_g2(pk->b2, w, pk->b_map, pk->plan_b, pk->msm[3], vslot(pk->results, 4), false, sB);
  if (!rc) FB_CUDA(fetch(4, sizeof(G2XYZZ) * MSM_VBITS, sB, 4));
  if (!rc) rc = msm_g1(pk->b1, w, pk->b_map, pk->plan_b, pk->msm[3], (G1XYZZ*)vslot(pk->results, 3), true, sB);
  if (!rc) FB_CUDA(fetch(3, sizeof(G1XYZZ) * MSM_VBITS, sB, 3));
  if (!rc) rc = msm_g1(pk->a, w, pk->a_map, pk->plan_a, pk->msm[2], (G1XYZZ*)vslot(pk->results, 2), false, sA);
  if (!rc) FB_CUDA(fetch(2, sizeof(G1XYZZ) * MSM_VBITS, sA, 2));
  if (!rc) rc = msm_g1(pk->l, w + pk->n_in + l_lo, nullptr, pk->plan_l, pk->msm[1], (G1XYZZ*)vslot(pk->results, 1), false, sL);
  if (!rc) FB_CUDA(fetch(1, sizeof(G1XYZZ) * MSM_VBITS, sL, 1));
  // R1CS evaluation and the H pipeline on the main stream
  if (rc) {
    set_error("MSM launch failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError()));
    return FB_ERR_CUDA;
  }
  if (pk->dist_g) {
    if (h_out) { set_error("h_out is not available on a distributed key"); return FB_ERR_ARG; }
    rc = eval_r1cs_cyclic(pk->csr, w, pk->n_in, pk->n_gates_global, pk->dist_g, pk->shard, pk->ev[0], pk->ev[1],
                          pk->ev[2], m >> pk->dist_g, st);
    if (rc) return rc;
    cudaEventRecord(T.ev[2], st);
    rc = pk->dom.dist_h_pipeline(pk->ev, pk->xtmp, pk->dist_g, pk->shard, dist_exchange(ctx), st);
    if (rc) { if (rc != FB_ERR_CUDA) set_error("distributed H pipeline failed (%d)", rc); return FB_ERR_CUDA; }
  } else {
    rc = eval_r1cs(pk->csr, w, pk->n_in, pk->ev[0], pk->ev[1], pk->ev[2], m, st);
    if (rc) return rc;
    cudaEventRecord(T.ev[2], st);
    for (int i = 0; i < 3; i++) pk->dom.ifft_then_coset_fft(pk->ev[i], st);
    pk->dom.pointwise_then_icoset_fft(pk->ev[0], pk->ev[1], pk->ev[2], st);
  }
  cudaEventRecord(T.ev[3], st);
  if (h_out) {
    pk->dom.bitrev(pk->scratch, pk->ev[0], st);
    FB_CUDA(cudaMemcpyAsync(h_out, pk->scratch, (m - 1) * sizeof(Fr), cudaMemcpyDeviceToHost, st));
  }
  rc = msm_g1(pk->h, pk->ev[0] + (pk->dist_g ? 0 : h_lo), nullptr, pk->plan_h, pk->msm[0],
              (G1XYZZ*)vslot(pk->results, 0), false, st);
  if (rc) { set_error("MSM launch failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError())); return FB_ERR_CUDA; }
  FB_CUDA(fetch(0, sizeof(G1XYZZ) * MSM_VBITS, st, 0));
  cudaEventRecord(T.ev[4], st);
  return FB_OK;
}

static void collect_timings(double host_ms, double total_ms) {
  Timing& T = g_timing;
  cudaEventElapsedTime(&T.ms[0], T.ev[0], T.ev[1]);
  cudaEventElapsedTime(&T.ms[1], T.ev[1], T.ev[2]);
  cudaEventElapsedTime(&T.ms[2], T.ev[2], T.ev[3]);
  cudaEventElapsedTime(&T.ms[3], T.ev[3], T.ev[4]);
  T.ms[4] = (float)host_ms;
  T.ms[5] = (float)total_ms;
}

// Small keys (domain <= 2^16): a prove is ~70 launches of tiny kernels on four streams and is bound by
// launch latency, not by the GPU.  The device side is captured once per key (per batch slot) as a CUDA
// graph -- witness upload from a pinned staging buffer, the witness-only MSMs forked onto their streams,
// R1CS + H pipeline + H MSM, the five result copies, joined back -- and replayed with one launch per
// proof; the host tails then run on the calling thread (a few microseconds each at these window sizes).
static int prove_graph(Ctx* ctx, ProvingKey* pk, const uint64_t* inputs, uint32_t n_in, const uint64_t* aux,
                       uint32_t n_aux, const uint64_t* r, const uint64_t* s, uint8_t* proof_raw, bool* fell_back) {
  *fell_back = false;
  auto t0 = std::chrono::steady_clock::now();
  FB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t wbytes = (size_t)(n_in + n_aux) * sizeof(Fr);
  if (!pk->w_stage) FB_CUDA(cudaMallocHost(&pk->w_stage, std::max<size_t>(wbytes, 32)));
  if (!pk->graph_exec) {
    const unsigned long long l0 = g_launches.load();
    bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    int rc = FB_OK;
    if (ok) {
      ok = cudaMemcpyAsync(pk->w, pk->w_stage, wbytes, cudaMemcpyHostToDevice, st) == cudaSuccess;
      if (ok) rc = prove_launch(pk, nullptr);
      for (int e = 3; e >= 1 && ok && !rc; e--) ok = cudaStreamWaitEvent(st, pk->msm_done[e], 0) == cudaSuccess;
      cudaGraph_t g = nullptr;
      const cudaError_t ee = cudaStreamEndCapture(st, &g);
      ok = ok && !rc && ee == cudaSuccess && g;
      if (ok) ok = cudaGraphInstantiate(&pk->graph_exec, g, 0) == cudaSuccess;
      if (g) cudaGraphDestroy(g);
    }
    // kernels enqueued between the two reads by THIS thread were captured, not run: take them off again (a
    // subtraction, so launches counted meanwhile by other batch slots are kept; their capture windows may
    // inflate graph_launches of this key slightly when keys are captured concurrently -- instrumentation only)
    pk->graph_launches = g_launches.load() - l0;
    g_launches.fetch_sub(pk->graph_launches, std::memory_order_relaxed);
    if (!ok) {
      cudaGetLastError();
      pk->graph_exec = nullptr;
      pk->graph_failed = true;
      *fell_back = true;
      return FB_OK;
    }
  }
  memcpy(pk->w_stage, inputs, (size_t)n_in * sizeof(Fr));
  if (n_aux) memcpy((uint8_t*)pk->w_stage + (size_t)n_in * sizeof(Fr), aux, (size_t)n_aux * sizeof(Fr));
  FB_CUDA(cudaGraphLaunch(pk->graph_exec, st));
  count_launch((int)pk->graph_launches);
  FixedTerms ft;
  VkPoints vk{pk->alpha_g1, pk->beta_g1, pk->delta_g1, pk->beta_g2, pk->delta_g2};
  const int rc_fixed = fixed_terms(&vk, reinterpret_cast<const KeyTables*>(pk->host_tables), r, s, ft);
  FB_CUDA(cudaStreamSynchronize(st));
  FB_CUDA(cudaGetLastError());
  if (rc_fixed) return rc_fixed;
  auto t1 = std::chrono::steady_clock::now();
  const H2 B2 = msm_horner_host<HFq2>(vslot(pk->results_host, 4), pk->plan_b.vbits());
  const H1 B1 = msm_horner_host<HFq>((const G1XYZZ*)vslot(pk->results_host, 3), pk->plan_b.vbits());
  const H1 A = msm_horner_host<HFq>((const G1XYZZ*)vslot(pk->results_host, 2), pk->plan_a.vbits());
  const H1 L = msm_horner_host<HFq>((const G1XYZZ*)vslot(pk->results_host, 1), pk->plan_l.vbits());
  const H1 H = msm_horner_host<HFq>((const G1XYZZ*)vslot(pk->results_host, 0), pk->plan_h.vbits());
  // the two 254-bit scalar multiplications of the assembly (s * A, r * B1: ~0.1 ms each) side by side
  auto fut_rb1 = std::async(std::launch::async, [&]() -> H1 { return scalar_mul(B1, ft.rc); });
  const H1 sA = scalar_mul(A, ft.sc);
  const H1 rB1 = fut_rb1.get();
  finish_proof(ft, H, L, A, B1, B2, &sA, &rB1, proof_raw);
  auto t2 = std::chrono::steady_clock::now();
  Timing& T = g_timing;  // stage events cannot be read out of a graph: only host and total are reported
  T.ms[0] = T.ms[1] = T.ms[2] = T.ms[3] = 0;
  T.ms[4] = (float)std::chrono::duration<double, std::milli>(t2 - t1).count();
  T.ms[5] = (float)std::chrono::duration<double, std::milli>(t2 - t0).count();
  return FB_OK;
}

// Host -> device copy of a witness range on stream `st`.  The reference hands over `Vec<Num<Fr>>` (cs.rs:99-102):
// PAGEABLE memory, which cudaMemcpyAsync moves through the driver's single staging buffer at a fraction of the
// PCIe rate.  Large pageable ranges are therefore staged here: four host threads memcpy 8 MiB chunks into a ring
// of pinned slots (two per thread) and queue one asynchronous DMA per chunk on their own copy streams, so the
// page-touching memcpys of one chunk overlap the DMA of the others; `st` then waits for the four streams.
// Pinned / registered memory (cudaHostAlloc, cudaHostRegister) and small ranges take the direct path.
static constexpr size_t kStageChunk = 8u << 20;
static constexpr int kStageThreads = 4, kStageSlots = 8;
static int upload_host(Ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return FB_OK;
  bool pageable = false;
  if (bytes >= (4u << 20) && !getenv("FB_NO_STAGED_UPLOAD")) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, src) == cudaSuccess) pageable = at.type == cudaMemoryTypeUnregistered;
    else cudaGetLastError();
  }
  if (!pageable) {
    FB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return FB_OK;
  }
  if (!ctx->stage) {
    FB_CUDA(cudaMallocHost(&ctx->stage, kStageChunk * kStageSlots));
    for (auto& s : ctx->copy_stream) FB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    for (auto& e : ctx->stage_ev) FB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : ctx->copy_done) FB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  // the copy streams must not run ahead of work already queued on st that still reads the destination
  cudaEvent_t start = ctx->copy_done[0];
  FB_CUDA(cudaEventRecord(start, st));
  const size_t nchunks = (bytes + kStageChunk - 1) / kStageChunk;
  std::atomic<int> err{0};
  std::vector<std::thread> th;
  for (int t = 0; t < kStageThreads; t++) {
    th.emplace_back([&, t] {
      cudaSetDevice(ctx->device);
      cudaStream_t cs = ctx->copy_stream[t];
      if (cudaStreamWaitEvent(cs, start, 0) != cudaSuccess) err = 1;
      bool used[2] = {false, false};
      int k = 0;
      for (size_t c = t; c < nchunks && !err; c += kStageThreads, k ^= 1) {
        const int slot = t * 2 + k;
        uint8_t* sb = reinterpret_cast<uint8_t*>(ctx->stage) + (size_t)slot * kStageChunk;
        if (used[k] && cudaEventSynchronize(ctx->stage_ev[slot]) != cudaSuccess) err = 1;
        const size_t off = c * kStageChunk, len = std::min(kStageChunk, bytes - off);
        memcpy(sb, reinterpret_cast<const uint8_t*>(src) + off, len);
        if (cudaMemcpyAsync(reinterpret_cast<uint8_t*>(dst) + off, sb, len, cudaMemcpyHostToDevice, cs) != cudaSuccess) err = 1;
        if (cudaEventRecord(ctx->stage_ev[slot], cs) != cudaSuccess) err = 1;
        used[k] = true;
      }
    });
  }
  for (auto& x : th) x.join();
  if (err) { set_error("staged witness upload failed: %s", cudaGetErrorString(cudaGetLastError())); return FB_ERR_CUDA; }
  for (int t = 0; t < kStageThreads; t++) {
    FB_CUDA(cudaEventRecord(ctx->copy_done[t], ctx->copy_stream[t]));
    FB_CUDA(cudaStreamWaitEvent(st, ctx->copy_done[t], 0));
  }
  // the ring is reused by the next call: its last DMAs must have left the slots before this returns to a caller
  // that may immediately upload again -- stage_ev are synchronised on reuse above, and across calls here
  for (int t = 0; t < kStageThreads; t++) FB_CUDA(cudaStreamSynchronize(ctx->copy_stream[t]));
  return FB_OK;
}

static int prove_impl(Ctx* ctx, ProvingKey* pk, const uint64_t* inputs, uint32_t n_in,
                      const uint64_t* aux, uint32_t n_aux, const void* dev_w, const uint64_t* r,
                      const uint64_t* s, uint8_t* proof_raw, uint8_t* partial, uint64_t* h_out) {
  if (!ctx || !pk) { set_error("fb_prove: null handle"); return FB_ERR_ARG; }
  if (ctx != pk->ctx) {  // the key's arrays, streams and events live on the context it was loaded on
    set_error("fb_prove: the key was loaded on a different fb_ctx");
    return FB_ERR_ARG;
  }
  if (!dev_w && (n_in != pk->n_in || n_aux != pk->n_aux)) {
    set_error("witness has n_in=%u n_aux=%u, key expects %u / %u", n_in, n_aux, pk->n_in, pk->n_aux);
    return FB_ERR_ARG;
  }
  static const int graph_env = [] { const char* e = getenv("FB_PROVE_GRAPH"); return (e && atoi(e) == 0) ? 0 : 1; }();
  const int use_graph = g_prove_graph.load(std::memory_order_relaxed) < 0 ? graph_env : g_prove_graph.load(std::memory_order_relaxed);
  if (use_graph && !dev_w && !partial && !h_out && !g_serial && !kstat_enabled() && pk->nshards == 1 &&
      !pk->dist_g && pk->m <= (1ull << 16) && !pk->graph_failed) {
    if (!g_timing.init) {
      for (auto& e : g_timing.ev) cudaEventCreate(&e);
      g_timing.init = true;
    }
    bool fell_back = false;
    const int grc = prove_graph(ctx, pk, inputs, n_in, aux, n_aux, r, s, proof_raw, &fell_back);
    if (!fell_back) return grc;
  }
  auto t0 = std::chrono::steady_clock::now();
  FB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (!g_timing.init) {
    for (auto& e : g_timing.ev) cudaEventCreate(&e);
    g_timing.init = true;
  }
  cudaEventRecord(g_timing.ev[0], st);
  if (dev_w) {
    // the witness is already in HBM: every kernel of the prove reads the caller's buffer in place (the call is
    // synchronous, so the buffer outlives them); a 512 MiB device-to-device copy would cost 0.3 ms per prove
  } else if (ctx->exchange && ctx->world == pk->nshards && pk->nshards > 1 && ctx->rank == pk->shard) {
    // every rank uploads 1/world of the witness over its own PCIe link, NVLink all-gathers the rest
    const uint64_t total = (uint64_t)n_in + n_aux, W = pk->nshards;
    const uint64_t chunk = (total + W - 1) / W;
    const uint64_t lo = chunk * pk->shard, hi = std::min(total, lo + chunk);
    for (uint64_t pos = lo; pos < hi;) {  // [lo, hi) may straddle the inputs | aux boundary
      const bool in_inputs = pos < n_in;
      const uint64_t seg_end = in_inputs ? std::min<uint64_t>(hi, n_in) : hi;
      const uint64_t* src = in_inputs ? inputs + 4 * pos : aux + 4 * (pos - n_in);
      { int urc = upload_host(ctx, pk->w + pos, src, (seg_end - pos) * sizeof(Fr), st); if (urc) return urc; }
      pos = seg_end;
    }
    int grc = dist_all_gather_inplace(ctx, pk->w, chunk * sizeof(Fr), st);
    if (grc) return grc;
  } else {
    int urc = upload_host(ctx, pk->w, inputs, (size_t)n_in * sizeof(Fr), st);
    if (!urc) urc = upload_host(ctx, pk->w + n_in, aux, (size_t)n_aux * sizeof(Fr), st);
    if (urc) return urc;
  }
  auto tl0 = std::chrono::steady_clock::now();
  int rc = prove_launch(pk, h_out, reinterpret_cast<const Fr*>(dev_w));
  if (rc) { cudaDeviceSynchronize(); return rc; }
  auto tl1 = std::chrono::steady_clock::now();
  // host work that needs only r, s and the key overlaps the device
  FixedTerms ft;
  VkPoints vk{pk->alpha_g1, pk->beta_g1, pk->delta_g1, pk->beta_g2, pk->delta_g2};
  int rc_fixed = partial ? FB_OK : fixed_terms(&vk, reinterpret_cast<const KeyTables*>(pk->host_tables), r, s, ft);
  // finish each MSM on the host as soon as its bit sums arrive (B2, B1, A, L finish while the
  // device still runs the H pipeline and the H MSM); the three heaviest tails get their own threads
  const int dev = ctx->device;
  const bool full = !partial && !rc_fixed;
  auto fut_b2 = std::async(std::launch::async, [&, dev]() -> H2 {
    cudaSetDevice(dev);
    cudaEventSynchronize(pk->msm_done[4]);
    return msm_horner_host<HFq2>(vslot(pk->results_host, 4), pk->plan_b.vbits());
  });
  auto fut_b1 = std::async(std::launch::async, [&, dev]() -> std::pair<H1, H1> {
    cudaSetDevice(dev);
    cudaEventSynchronize(pk->msm_done[3]);
    H1 b1 = msm_horner_host<HFq>((const G1XYZZ*)vslot(pk->results_host, 3), pk->plan_b.vbits());
    return {b1, full ? scalar_mul(b1, ft.rc) : H1::inf()};
  });
  auto fut_a = std::async(std::launch::async, [&, dev]() -> std::pair<H1, H1> {
    cudaSetDevice(dev);
    cudaEventSynchronize(pk->msm_done[2]);
    H1 a = msm_horner_host<HFq>((const G1XYZZ*)vslot(pk->results_host, 2), pk->plan_a.vbits());
    return {a, full ? scalar_mul(a, ft.sc) : H1::inf()};
  });
  FB_CUDA(cudaEventSynchronize(pk->msm_done[1]));
  H1 L = msm_horner_host<HFq>((const G1XYZZ*)vslot(pk->results_host, 1), pk->plan_l.vbits());
  FB_CUDA(cudaEventSynchronize(pk->msm_done[0]));
  auto t1 = std::chrono::steady_clock::now();
  H1 H = msm_horner_host<HFq>((const G1XYZZ*)vslot(pk->results_host, 0), pk->plan_h.vbits());
  H2 B2 = fut_b2.get();
  std::pair<H1, H1> pb1 = fut_b1.get(), pa = fut_a.get();
  const H1 &B1 = pb1.first, &A = pa.first;
  if (getenv("FB_TRACE")) {
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fprintf(stderr, "[fb trace] pre-launch %.3f  enqueue %.3f  fixed+waits %.3f  horner(H)+joins %.3f ms\n",
            ms(t0, tl0), ms(tl0, tl1), ms(tl1, t1), ms(t1, std::chrono::steady_clock::now()));
  }
  FB_CUDA(cudaStreamSynchronize(st));
  FB_CUDA(cudaGetLastError());
  if (rc_fixed) return rc_fixed;
  if (partial) {
    memset(partial, 0, 640);
    G1Affine p;
    p = h1_affine(H); memcpy(partial + 0, &p, 64);
    p = h1_affine(L); memcpy(partial + 128, &p, 64);
    p = h1_affine(A); memcpy(partial + 256, &p, 64);
    p = h1_affine(B1); memcpy(partial + 384, &p, 64);
    G2Affine q = h2_affine(B2);
    memcpy(partial + 512, &q, 128);
    rc = FB_OK;
  } else {
    finish_proof(ft, H, L, A, B1, B2, &pa.second, &pb1.second, proof_raw);
    rc = FB_OK;
  }
  auto t2 = std::chrono::steady_clock::now();
  collect_timings(std::chrono::duration<double, std::milli>(t2 - t1).count(),
                  std::chrono::duration<double, std::milli>(t2 - t0).count());
  return rc;
}

}  // namespace fb

using namespace fb;

extern "C" {

const char* fb_last_error(void) { return fb::last_error_cstr(); }

int fb_device_count(void) try {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
} FB_ABI_CATCH_INT

int fb_init(const int* devices, int ndev, fb_ctx** out) try {
  if (!out || ndev != 1) {
    set_error("fb_init: exactly one device per context (one process per GPU)");
    return FB_ERR_ARG;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device available (%s); this backend has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return FB_ERR_CUDA;
  }
  int dev = devices ? devices[0] : 0;
  if (dev < 0 || dev >= n) { set_error("device %d out of range (have %d)", dev, n); return FB_ERR_ARG; }
  FB_CUDA(cudaSetDevice(dev));
  if (const char* e = getenv("FB_L2_FETCH_BYTES")) {  // device-wide L2 fetch granularity hint (A/B measurements)
    size_t before = 0, after = 0;
    cudaDeviceGetLimit(&before, cudaLimitMaxL2FetchGranularity);
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e));
    cudaDeviceGetLimit(&after, cudaLimitMaxL2FetchGranularity);
    fprintf(stderr, "[fawkes_b200] L2 fetch granularity %zu -> %zu bytes\n", before, after);
  }
  Ctx* c = new Ctx();
  c->device = dev;
  FB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (int i = 0; i < 3; i++) {
    FB_CUDA(cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking));
    FB_CUDA(cudaEventCreateWithFlags(&c->aux_done[i], cudaEventDisableTiming));
  }
  *out = reinterpret_cast<fb_ctx*>(c);
  return FB_OK;
} FB_ABI_CATCH_INT

void fb_shutdown(fb_ctx* ctx) try {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return;
  cudaSetDevice(c->device);
  dist_destroy(c);  // NCCL communicator and the combine buffers, if fb_dist_init was called
  if (c->stage) {
    cudaFreeHost(c->stage);
    for (auto& s : c->copy_stream) cudaStreamDestroy(s);
    for (auto& e : c->stage_ev) cudaEventDestroy(e);
    for (auto& e : c->copy_done) cudaEventDestroy(e);
  }
  cudaStreamDestroy(c->stream);
  for (int i = 0; i < 3; i++) { cudaStreamDestroy(c->aux[i]); cudaEventDestroy(c->aux_done[i]); }
  delete c;
} FB_ABI_CATCH_VOID

void fb_free(void* p) { free(p); }

int fb_circuit_from_raw_gates(const uint8_t* gates, size_t len, uint32_t num_gates, uint32_t n_in,
                              uint32_t n_aux, fb_circuit** out) try {
  if (!out || (!gates && len)) { set_error("fb_circuit_from_raw_gates: bad argument"); return FB_ERR_ARG; }
  std::unique_ptr<Circuit> c(new Circuit());  // freed on every error path, exceptions included
  c->n_in = n_in;
  c->n_aux = n_aux;
  int rc = parse_gates_to_csr(gates, len, n_in, n_aux, c->csr);
  if (rc) return rc;
  if (c->csr.n_gates != num_gates) {
    set_error("gate stream holds %u gates, Parameters announce %u", c->csr.n_gates, num_gates);
    return FB_ERR_FORMAT;
  }
  *out = reinterpret_cast<fb_circuit*>(c.release());
  return FB_OK;
} FB_ABI_CATCH_INT

// Most bytes a gate stream announcing `num_gates` gates over n_in + n_aux variables can hold: three LCs per
// gate, each a u32 count and at most one 37-byte term per variable (LCs are kept canonical, lc.rs:89-118);
// capped by an absolute limit (FB_MAX_GATE_BYTES, default 32 GiB: configs[4] is ~9 GB).
static size_t gate_stream_cap(uint32_t num_gates, uint32_t n_in, uint32_t n_aux) {
  size_t abs_cap = (size_t)32 << 30;
  if (const char* e = getenv("FB_MAX_GATE_BYTES")) abs_cap = (size_t)strtoull(e, nullptr, 10);
  const long double theo = (long double)num_gates * 3.0L * (4.0L + 37.0L * ((long double)n_in + n_aux));
  return theo < (long double)abs_cap ? (size_t)theo + 1 : abs_cap;
}

int fb_circuit_from_gates(const uint8_t* gates_brotli, size_t len, uint32_t num_gates, uint32_t n_in,
                          uint32_t n_aux, fb_circuit** out) try {
  std::vector<uint8_t> raw;
  int rc = brotli_decode(gates_brotli, len, raw, gate_stream_cap(num_gates, n_in, n_aux));
  if (rc) return rc;
  return fb_circuit_from_raw_gates(raw.data(), raw.size(), num_gates, n_in, n_aux, out);
} FB_ABI_CATCH_INT

int fb_circuit_from_raw_gates_gpu(fb_ctx* ctx, const uint8_t* gates, size_t len, uint32_t num_gates, uint32_t n_in,
                                  uint32_t n_aux, fb_circuit** out, float* times_ms) try {
  if (!ctx || !out || (!gates && len)) { set_error("fb_circuit_from_raw_gates_gpu: bad argument"); return FB_ERR_ARG; }
  std::unique_ptr<Circuit> c(new Circuit());  // freed on every error path, exceptions included
  c->n_in = n_in;
  c->n_aux = n_aux;
  int rc = parse_gates_device(reinterpret_cast<Ctx*>(ctx), gates, len, n_in, n_aux, c->csr, times_ms);
  if (rc) return rc;
  if (c->csr.n_gates != num_gates) {
    set_error("gate stream holds %u gates, Parameters announce %u", c->csr.n_gates, num_gates);
    return FB_ERR_FORMAT;
  }
  *out = reinterpret_cast<fb_circuit*>(c.release());
  return FB_OK;
} FB_ABI_CATCH_INT

int fb_circuit_from_gates_gpu(fb_ctx* ctx, const uint8_t* gates_brotli, size_t len, uint32_t num_gates, uint32_t n_in,
                              uint32_t n_aux, fb_circuit** out, float* times_ms) try {
  std::vector<uint8_t> raw;
  auto t0 = std::chrono::steady_clock::now();
  int rc = brotli_decode(gates_brotli, len, raw, gate_stream_cap(num_gates, n_in, n_aux));
  if (rc) return rc;
  if (times_ms) times_ms[4] = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return fb_circuit_from_raw_gates_gpu(ctx, raw.data(), raw.size(), num_gates, n_in, n_aux, out, times_ms);
} FB_ABI_CATCH_INT

void fb_circuit_free(fb_circuit* c) { delete reinterpret_cast<Circuit*>(c); }

int fb_circuit_shape(const fb_circuit* c_, uint32_t* n_in, uint32_t* n_aux, uint32_t* n_gates,
                     uint64_t* nnz) try {
  const Circuit* c = reinterpret_cast<const Circuit*>(c_);
  if (!c) return FB_ERR_ARG;
  if (n_in) *n_in = c->n_in;
  if (n_aux) *n_aux = c->n_aux;
  if (n_gates) *n_gates = c->csr.n_gates;
  if (nnz) *nnz = c->csr.col[0].size() + c->csr.col[1].size() + c->csr.col[2].size();
  return FB_OK;
} FB_ABI_CATCH_INT

int fb_pk_load_shard(fb_ctx* ctx, const uint8_t* bellman_params, size_t len,
                     const fb_circuit* circuit, int checked, int shard, int nshards, fb_pk** out) try {
  return load_key(reinterpret_cast<Ctx*>(ctx), bellman_params, len,
                  reinterpret_cast<const Circuit*>(circuit), checked, shard, nshards,
                  reinterpret_cast<ProvingKey**>(out));
} FB_ABI_CATCH_INT

int fb_pk_load_circuit(fb_ctx* ctx, const uint8_t* bellman_params, size_t len,
                       const fb_circuit* circuit, int checked, fb_pk** out) try {
  return fb_pk_load_shard(ctx, bellman_params, len, circuit, checked, 0, 1, out);
} FB_ABI_CATCH_INT

int fb_pk_load(fb_ctx* ctx, const uint8_t* bellman_params, size_t len, const uint8_t* gates_brotli,
               size_t glen, uint32_t num_gates, int checked, fb_pk** out) try {
  ParamsView v;
  int rc = parse_params(bellman_params, len, v);
  if (rc) return rc;
  fb_circuit* c = nullptr;
  rc = fb_circuit_from_gates_gpu(ctx, gates_brotli, glen, num_gates, v.n_ic, v.n_l, &c, nullptr);
  if (rc) return rc;
  rc = fb_pk_load_circuit(ctx, bellman_params, len, c, checked, out);
  fb_circuit_free(c);
  return rc;
} FB_ABI_CATCH_INT

void fb_pk_free(fb_pk* pk) try {
  ProvingKey* p = reinterpret_cast<ProvingKey*>(pk);
  if (p && p->ctx) cudaSetDevice(p->ctx->device);
  pk_release(p);
} FB_ABI_CATCH_VOID

int fb_pk_get_info(const fb_pk* pk_, fb_pk_info* info) try {
  const ProvingKey* pk = reinterpret_cast<const ProvingKey*>(pk_);
  if (!pk || !info) return FB_ERR_ARG;
  info->n_in = pk->n_in;
  info->n_aux = pk->n_aux;
  info->n_gates = pk->csr.n_gates;
  info->log_m = pk->k;
  info->len_h = pk->len_h;
  info->len_l = pk->plan_l.n;
  info->len_a = pk->len_a;
  info->len_b = pk->len_b;
  info->nnz = pk->csr.nnz[0] + pk->csr.nnz[1] + pk->csr.nnz[2];
  uint64_t b = (uint64_t)pk->len_h * 64 + (uint64_t)pk->plan_l.n * 64 + (uint64_t)pk->len_a * 68 +
               (uint64_t)pk->len_b * (64 + 128 + 4) + info->nnz * 8 + 5 * pk->m * 32 +
               3 * (pk->m - 1) * 32 + (uint64_t)(pk->n_in + pk->n_aux) * 32;
  info->hbm_bytes = b + pk->table_bytes;
  info->g1_digit_slots = (uint64_t)pk->plan_h.n * pk->plan_h.W + (uint64_t)pk->plan_l.n * pk->plan_l.W +
                         (uint64_t)pk->plan_a.n * pk->plan_a.W + (uint64_t)pk->plan_b.n * pk->plan_b.W;
  info->g2_digit_slots = (uint64_t)pk->plan_b.n * pk->plan_b.W;
  info->msm_window_bits = pk->plan_h.c;
  info->msm_windows = pk->plan_h.W;
  info->msm_tables = pk->plan_h.table ? 1 : 0;
  info->reserved0 = 0;
  info->table_bytes = pk->table_bytes;
  return FB_OK;
} FB_ABI_CATCH_INT

// Sharded key on a distributed context (fb_dist_init + fb_pk_load_shard(rank, world)): a COLLECTIVE prove.  Every
// rank calls with the same witness, r and s; each computes the partial sums of its base shard, the 640-byte
// partials are all-gathered over the library's NCCL communicator (NVLink) and every rank assembles and returns
// the same proof.
static int prove_collective(Ctx* ctx, ProvingKey* p, const uint64_t* inputs, uint32_t n_in, const uint64_t* aux,
                            uint32_t n_aux, const void* dev_w, const uint64_t r[4], const uint64_t s[4],
                            uint8_t proof_raw[256]) {
  if (!ctx->exchange || ctx->world != p->nshards || ctx->rank != p->shard) {
    set_error("fb_prove on a sharded key needs a context initialised with fb_dist_init(rank = shard, world = nshards); "
              "use fb_prove_partial / fb_prove_finish with your own transport otherwise");
    return FB_ERR_ARG;
  }
  uint8_t mine[640];
  int rc = prove_impl(ctx, p, inputs, n_in, aux, n_aux, dev_w, nullptr, nullptr, nullptr, mine, nullptr);
  if (rc) return rc;
  auto t0 = std::chrono::steady_clock::now();
  std::vector<uint8_t> all((size_t)ctx->world * 640);
  rc = dist_gather_partials(ctx, mine, all.data(), ctx->stream);
  if (rc) return rc;
  H1 sum[4] = {H1::inf(), H1::inf(), H1::inf(), H1::inf()};
  H2 sum2 = H2::inf();
  for (int i = 0; i < ctx->world; i++) {
    const uint8_t* q = all.data() + (size_t)i * 640;
    for (int j = 0; j < 4; j++) {
      G1Affine a;
      memcpy(&a, q + 128 * j, 64);
      sum[j] = add(sum[j], h1_from(a));
    }
    G2Affine b;
    memcpy(&b, q + 512, 128);
    sum2 = add(sum2, h2_from(b));
  }
  FixedTerms ft;
  VkPoints vk{p->alpha_g1, p->beta_g1, p->delta_g1, p->beta_g2, p->delta_g2};
  rc = fixed_terms(&vk, reinterpret_cast<const KeyTables*>(p->host_tables), r, s, ft);
  if (rc) return rc;
  auto fut_rb1 = std::async(std::launch::async, [&]() -> H1 { return scalar_mul(sum[3], ft.rc); });
  const H1 sA = scalar_mul(sum[2], ft.sc);
  const H1 rB1 = fut_rb1.get();
  finish_proof(ft, sum[0], sum[1], sum[2], sum[3], sum2, &sA, &rB1, proof_raw);
  const double combine_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  g_timing.ms[4] += (float)combine_ms;
  g_timing.ms[5] += (float)combine_ms;
  return FB_OK;
}

int fb_prove(fb_ctx* ctx, fb_pk* pk, const uint64_t* inputs, uint32_t n_in, const uint64_t* aux,
             uint32_t n_aux, const uint64_t r[4], const uint64_t s[4], uint8_t proof_raw[256],
             uint64_t* h_out) try {
  if (!inputs || (!aux && n_aux) || !r || !s || !proof_raw) { set_error("fb_prove: null buffer"); return FB_ERR_ARG; }
  ProvingKey* p = reinterpret_cast<ProvingKey*>(pk);
  if (p && p->nshards != 1) {
    if (h_out) { set_error("h_out is not available on a sharded key"); return FB_ERR_ARG; }
    if (!ctx) { set_error("fb_prove: null handle"); return FB_ERR_ARG; }
    return prove_collective(reinterpret_cast<Ctx*>(ctx), p, inputs, n_in, aux, n_aux, nullptr, r, s, proof_raw);
  }
  return prove_impl(reinterpret_cast<Ctx*>(ctx), p, inputs, n_in, aux, n_aux, nullptr, r, s,
                    proof_raw, nullptr, h_out);
} FB_ABI_CATCH_INT

// A slot = everything one prove writes (witness, evaluations, MSM scratch, result buffers, events) plus its
// own streams, next to a borrowed view of the key's immutable arrays.
static ProvingKey* make_slot(const ProvingKey* pk) {
  ProvingKey* s = new ProvingKey(*pk);
  s->is_slot = true;
  s->slots.clear();
  s->graph_exec = nullptr; s->w_stage = nullptr; s->graph_launches = 0; s->graph_failed = false;
  s->batch = nullptr;
  s->w = nullptr; s->scratch = nullptr; s->results = nullptr; s->results_host = nullptr;
  for (auto& p : s->ev) p = nullptr;
  for (auto& p : s->xtmp) p = nullptr;
  for (auto& m : s->msm) m = MsmScratch();
  for (auto& e : s->msm_done) e = nullptr;
  Ctx* c = new Ctx();
  c->device = pk->ctx->device;
  s->ctx = c;
  bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < 3 && ok; i++)
    ok = cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->aux_done[i], cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaMalloc(&s->w, ((size_t)(pk->n_in + pk->n_aux) + 1) * sizeof(Fr)) == cudaSuccess;
  for (int i = 0; i < 3 && ok; i++) ok = cudaMalloc(&s->ev[i], pk->m * sizeof(Fr)) == cudaSuccess;
  ok = ok && cudaMalloc(&s->scratch, pk->m * sizeof(Fr)) == cudaSuccess;
  const MsmPlan plans[4] = {pk->plan_h, pk->plan_l, pk->plan_a, pk->plan_b};
  for (int i = 0; i < 4 && ok; i++) ok = s->msm[i].alloc(&plans[i], 1, i == 3) == 0;
  ok = ok && cudaMalloc(&s->results, 5 * MSM_VBITS * sizeof(G2XYZZ)) == cudaSuccess;
  for (auto& e : s->msm_done) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaMallocHost(&s->results_host, 5 * MSM_VBITS * sizeof(G2XYZZ)) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    pk_release(s);
    return nullptr;
  }
  return s;
}

// ---- batched proves of one small circuit (BASELINE configs[1]: 256 EdDSA proofs per run) ---------------------
// The reference proves them one prove() call at a time (prover.rs:63-90).  One such prove is ~70 tiny kernels and
// leaves a B200 idle, so a batch is proved as ONE set of launches carrying `P` proofs each: the witnesses are
// uploaded as one [P][n_in + n_aux] array, R1CS evaluation and the seven transforms run with the proof index as
// grid.y, and each of the five MSMs is one batched MSM -- the key's bases are shared, the buckets are keyed by
// (proof, digit), so one digit sort, one bucket accumulation and one reduction serve all P proofs (msm.cuh:
// MsmPlan::batched).  Host tails (Horner over the bit sums, r/s terms, affine conversion: ~0.4 ms per proof) run
// on a pool of host threads for chunk k while the GPU works on chunk k + 1.
}  // extern "C"

namespace fb {
struct BatchWork {
  uint32_t P = 0;
  uint64_t wstride = 0;
  Fr* w[2] = {nullptr, nullptr};   // double buffered: the witness-only MSMs of chunk k+1 run beside chunk k's H pipeline
  Fr* ev[3] = {nullptr, nullptr, nullptr};
  MsmScratch msm[4];
  size_t vmax = 0;                 // V entries per proof per MSM (max over the plans)
  void* results = nullptr;         // device: 5 x P x vmax G2XYZZ-sized slots
  void* results_host[2] = {nullptr, nullptr};
  Fr* w_stage[2] = {nullptr, nullptr};
  cudaEvent_t done[2][5];
  cudaEvent_t uploaded[2];
  cudaEvent_t r1cs_done[2];        // the main stream has finished reading w[q]
  cudaStream_t up = nullptr;       // witness uploads
  bool used[2] = {false, false};
  bool ok = false;
  void release() {
    for (auto& p : w) cudaFree(p);
    for (auto& p : ev) cudaFree(p);
    for (auto& m : msm) m.release();
    cudaFree(results);
    for (auto& p : results_host) if (p) cudaFreeHost(p);
    for (auto& p : w_stage) if (p) cudaFreeHost(p);
    if (ok) {
      for (auto& q : done) for (auto& e : q) cudaEventDestroy(e);
      for (auto& e : uploaded) cudaEventDestroy(e);
      for (auto& e : r1cs_done) cudaEventDestroy(e);
      cudaStreamDestroy(up);
    }
  }
};
static void free_batch(void* p) {
  BatchWork* b = reinterpret_cast<BatchWork*>(p);
  if (b) { b->release(); delete b; }
}

static BatchWork* make_batch(ProvingKey* pk, uint32_t P) {
  BatchWork* b = new BatchWork();
  b->P = P;
  b->wstride = (uint64_t)pk->n_in + pk->n_aux;
  const MsmPlan plans[4] = {pk->plan_h.batched(P, pk->m), pk->plan_l.batched(P, b->wstride), pk->plan_a.batched(P, b->wstride),
                            pk->plan_b.batched(P, b->wstride)};
  for (const MsmPlan& p : plans) b->vmax = std::max<size_t>(b->vmax, (size_t)p.vbits_per_set());
  bool ok = true;
  for (int q = 0; q < 2 && ok; q++) ok = cudaMalloc(&b->w[q], std::max<size_t>(P * b->wstride, 1) * sizeof(Fr)) == cudaSuccess;
  for (int i = 0; i < 3 && ok; i++) ok = cudaMalloc(&b->ev[i], (size_t)P * pk->m * sizeof(Fr)) == cudaSuccess;
  for (int i = 0; i < 4 && ok; i++) ok = b->msm[i].alloc(&plans[i], 1, i == 3) == 0;
  const size_t rbytes = 5 * (size_t)P * b->vmax * sizeof(G2XYZZ);
  ok = ok && cudaMalloc(&b->results, rbytes) == cudaSuccess;
  for (int q = 0; q < 2 && ok; q++) {
    ok = cudaMallocHost(&b->results_host[q], rbytes) == cudaSuccess &&
         cudaMallocHost(&b->w_stage[q], std::max<size_t>(P * b->wstride, 1) * sizeof(Fr)) == cudaSuccess;
  }
  if (ok) {
    for (auto& q : b->done) for (auto& e : q) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    for (auto& e : b->uploaded) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    for (auto& e : b->r1cs_done) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&b->up, cudaStreamNonBlocking) == cudaSuccess;
    b->ok = true;
  }
  if (!ok) {
    cudaGetLastError();
    free_batch(b);
    return nullptr;
  }
  return b;
}

static int prove_batched(Ctx* ctx, ProvingKey* pk, uint32_t count, const uint64_t* const* inputs, const uint64_t* const* aux,
                         const uint64_t* r, const uint64_t* s, uint8_t* proofs_raw) {
  FB_CUDA(cudaSetDevice(ctx->device));
  uint32_t P = 64;  // measured on configs[1] (256 proofs, c = 10): 64 -> 0.157, 128 -> 0.163, 256 -> 0.177 ms per proof
  if (const char* e = getenv("FB_BATCH_P")) P = (uint32_t)std::max(1, std::min(1024, atoi(e)));
  // the workspaces are sized for a full chunk whatever this call's count is: the streaming worker proves chunks of
  // growing size, and re-allocating them per call cost more than the proofs
  BatchWork* bw = reinterpret_cast<BatchWork*>(pk->batch);
  if (bw && bw->P != P) { free_batch(bw); bw = nullptr; pk->batch = nullptr; }
  if (!bw) {
    bw = make_batch(pk, P);
    if (!bw) { set_error("fb_prove_batch: cannot allocate the workspaces of a %u-proof batch", P); return FB_ERR_CUDA; }
    pk->batch = bw;
  }
  P = std::min(bw->P, count);
  cudaStream_t st = ctx->stream, sL = ctx->aux[0], sA = ctx->aux[1], sB = ctx->aux[2];
  const uint64_t m = pk->m, ws = bw->wstride;
  const size_t slot_elems = (size_t)P * bw->vmax;   // G2XYZZ-sized elements per MSM slot
  auto dslot = [&](int i) { return reinterpret_cast<G2XYZZ*>(bw->results) + (size_t)i * slot_elems; };
  auto hslot = [&](int q, int i) { return reinterpret_cast<G2XYZZ*>(bw->results_host[q]) + (size_t)i * slot_elems; };
  VkPoints vk{pk->alpha_g1, pk->beta_g1, pk->delta_g1, pk->beta_g2, pk->delta_g2};
  const KeyTables* kt = reinterpret_cast<const KeyTables*>(pk->host_tables);
  const uint32_t nchunks = (count + P - 1) / P;
  int host_threads = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
  if (const char* e = getenv("FB_BATCH_HOST_THREADS")) host_threads = std::max(1, atoi(e));

  auto enqueue = [&](uint32_t k) -> int {
    const int q = (int)(k & 1);
    const uint32_t lo = k * P, cnt = std::min(P, count - lo);
    if (bw->used[q]) FB_CUDA(cudaEventSynchronize(bw->uploaded[q]));   // the stage buffer's last upload has left it
    for (uint32_t p = 0; p < cnt; p++) {
      Fr* dst = bw->w_stage[q] + (size_t)p * ws;
      memcpy(dst, inputs[lo + p], (size_t)pk->n_in * sizeof(Fr));
      if (pk->n_aux) memcpy(dst + pk->n_in, aux[lo + p], (size_t)pk->n_aux * sizeof(Fr));
    }
    // w[q] was last read by chunk k-2: its witness-only MSMs on the side streams and its R1CS evaluation
    Fr* w = bw->w[q];
    if (k >= 2) {
      for (int e = 1; e <= 4; e++) FB_CUDA(cudaStreamWaitEvent(bw->up, bw->done[q][e], 0));
      FB_CUDA(cudaStreamWaitEvent(bw->up, bw->r1cs_done[q], 0));
    }
    FB_CUDA(cudaMemcpyAsync(w, bw->w_stage[q], (size_t)cnt * ws * sizeof(Fr), cudaMemcpyHostToDevice, bw->up));
    FB_CUDA(cudaEventRecord(bw->uploaded[q], bw->up));
    bw->used[q] = true;
    for (int i = 0; i < 3; i++) FB_CUDA(cudaStreamWaitEvent(ctx->aux[i], bw->uploaded[q], 0));
    FB_CUDA(cudaStreamWaitEvent(st, bw->uploaded[q], 0));
    const MsmPlan ph = pk->plan_h.batched(cnt, m), pl = pk->plan_l.batched(cnt, ws), pa = pk->plan_a.batched(cnt, ws),
                  pb = pk->plan_b.batched(cnt, ws);
    auto fetch = [&](int slot, size_t bytes, cudaStream_t sx) -> cudaError_t {
      cudaError_t e = cudaMemcpyAsync(hslot(q, slot), dslot(slot), bytes, cudaMemcpyDeviceToHost, sx);
      if (e == cudaSuccess) e = cudaEventRecord(bw->done[q][slot], sx);
      return e;
    };
    int rc = msm_g2(pk->b2, w, pk->b_map, pb, bw->msm[3], dslot(4), false, sB);
    if (!rc) FB_CUDA(fetch(4, sizeof(G2XYZZ) * pb.vbits(), sB));
    if (!rc) rc = msm_g1(pk->b1, w, pk->b_map, pb, bw->msm[3], (G1XYZZ*)dslot(3), true, sB);
    if (!rc) FB_CUDA(fetch(3, sizeof(G1XYZZ) * pb.vbits(), sB));
    if (!rc) rc = msm_g1(pk->a, w, pk->a_map, pa, bw->msm[2], (G1XYZZ*)dslot(2), false, sA);
    if (!rc) FB_CUDA(fetch(2, sizeof(G1XYZZ) * pa.vbits(), sA));
    if (!rc) rc = msm_g1(pk->l, w + pk->n_in, nullptr, pl, bw->msm[1], (G1XYZZ*)dslot(1), false, sL);
    if (!rc) FB_CUDA(fetch(1, sizeof(G1XYZZ) * pl.vbits(), sL));
    if (rc) { set_error("batched MSM launch failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError())); return FB_ERR_CUDA; }
    rc = eval_r1cs_batch(pk->csr, w, ws, pk->n_in, bw->ev[0], bw->ev[1], bw->ev[2], m, cnt, st);
    if (rc) return rc;
    FB_CUDA(cudaEventRecord(bw->r1cs_done[q], st));
    for (int i = 0; i < 3; i++) pk->dom.ifft_then_coset_fft(bw->ev[i], st, cnt, m);
    pk->dom.pointwise_then_icoset_fft(bw->ev[0], bw->ev[1], bw->ev[2], st, cnt, m);
    rc = msm_g1(pk->h, bw->ev[0], nullptr, ph, bw->msm[0], (G1XYZZ*)dslot(0), false, st);
    if (rc) { set_error("batched MSM launch failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError())); return FB_ERR_CUDA; }
    FB_CUDA(fetch(0, sizeof(G1XYZZ) * ph.vbits(), st));
    return FB_OK;
  };

  auto tails = [&](uint32_t k) -> int {
    const int q = (int)(k & 1);
    const uint32_t lo = k * P, cnt = std::min(P, count - lo);
    for (int e = 0; e < 5; e++) FB_CUDA(cudaEventSynchronize(bw->done[q][e]));
    const int vh = pk->plan_h.vbits_per_set(), vl = pk->plan_l.vbits_per_set(), va = pk->plan_a.vbits_per_set(),
              vb = pk->plan_b.vbits_per_set();
    std::atomic<uint32_t> next{0};
    std::atomic<int> err{FB_OK};
    std::string err_msg;
    std::mutex err_mu;
    auto work = [&] {
      for (;;) {
        const uint32_t p = next.fetch_add(1);
        if (p >= cnt) return;
        const H2 B2 = msm_horner_host<HFq2>(hslot(q, 4) + (size_t)p * vb, vb);
        const H1 B1 = msm_horner_host<HFq>((const G1XYZZ*)hslot(q, 3) + (size_t)p * vb, vb);
        const H1 A = msm_horner_host<HFq>((const G1XYZZ*)hslot(q, 2) + (size_t)p * va, va);
        const H1 L = msm_horner_host<HFq>((const G1XYZZ*)hslot(q, 1) + (size_t)p * vl, vl);
        const H1 H = msm_horner_host<HFq>((const G1XYZZ*)hslot(q, 0) + (size_t)p * vh, vh);
        FixedTerms ft;
        const int frc = fixed_terms(&vk, kt, r + 4 * (size_t)(lo + p), s + 4 * (size_t)(lo + p), ft);
        if (frc) {
          std::lock_guard<std::mutex> lk(err_mu);
          err = frc;
          err_msg = last_error_cstr();
          return;
        }
        finish_proof(ft, H, L, A, B1, B2, nullptr, nullptr, proofs_raw + 256 * (size_t)(lo + p));
      }
    };
    const int T = std::min<int>(host_threads, (int)cnt);
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(work);
    work();
    for (auto& x : th) x.join();
    if (err) { set_error("%s", err_msg.c_str()); return err; }
    return FB_OK;
  };

  int rc = FB_OK;
  for (uint32_t k = 0; k < nchunks && !rc; k++) {
    rc = enqueue(k);
    if (!rc && k > 0) rc = tails(k - 1);
  }
  if (!rc) rc = tails(nchunks - 1);
  if (rc) cudaDeviceSynchronize();
  FB_CUDA(cudaGetLastError());
  return rc;
}
}  // namespace fb

extern "C" {

int fb_prove_batch(fb_ctx* ctx, fb_pk* pk_, uint32_t count, const uint64_t* const* inputs, uint32_t n_in,
                   const uint64_t* const* aux, uint32_t n_aux, const uint64_t* r, const uint64_t* s,
                   uint8_t* proofs_raw) try {
  if (!inputs || (!aux && n_aux) || !r || !s || !proofs_raw) { set_error("fb_prove_batch: null buffer"); return FB_ERR_ARG; }
  // The reference proves one circuit per prove() call (prover.rs:63-90); a batch is the same key used
  // `count` times.  Small keys (domain <= 2^16) go through prove_batched above: one set of launches per chunk of
  // proofs.  FB_BATCH_MODE=slots keeps the older scheme for A/B runs: up to FB_BATCH_SLOTS (default 8) independent
  // proves in flight, one host thread and one set of streams and workspaces per slot, all reading the same
  // resident key.  Big keys fill the GPU on their own and are proved one after the other.
  ProvingKey* pk = reinterpret_cast<ProvingKey*>(pk_);
  Ctx* c0 = reinterpret_cast<Ctx*>(ctx);
  if (!pk || !c0) { set_error("fb_prove_batch: null handle"); return FB_ERR_ARG; }
  if (n_in != pk->n_in || n_aux != pk->n_aux) {
    set_error("witness has n_in=%u n_aux=%u, key expects %u / %u", n_in, n_aux, pk->n_in, pk->n_aux);
    return FB_ERR_ARG;
  }
  if (c0 != pk->ctx) { set_error("fb_prove_batch: the key was loaded on a different fb_ctx"); return FB_ERR_ARG; }
  // small keys: one set of launches per chunk of proofs (prove_batched); FB_BATCH_MODE=slots keeps the older
  // scheme below (independent proves in flight on their own streams, CUDA-graph replay)
  {
    const char* mode = getenv("FB_BATCH_MODE");
    if (count >= 2 && pk->m <= (1u << 16) && pk->nshards == 1 && !pk->dist_g && !g_serial && !(mode && !strcmp(mode, "slots")))
      return prove_batched(c0, pk, count, inputs, aux, r, s, proofs_raw);
  }
  int want = 8;
  if (const char* e = getenv("FB_BATCH_SLOTS")) want = std::max(1, std::min(32, atoi(e)));
  // big keys saturate the GPU on their own (and their workspaces are large); sharded keys prove collectively
  if (pk->m > (1u << 16) || pk->nshards != 1 || pk->dist_g || g_serial) want = 1;
  want = (int)std::min<uint32_t>((uint32_t)want, count);
  FB_CUDA(cudaSetDevice(c0->device));
  while ((int)pk->slots.size() + 1 < want) {
    ProvingKey* sl = make_slot(pk);
    if (!sl) break;  // out of memory: run with the slots we have
    pk->slots.push_back(sl);
  }
  const int K = std::min<int>(want, (int)pk->slots.size() + 1);
  if (K <= 1) {
    for (uint32_t i = 0; i < count; i++) {
      int rc = fb_prove(ctx, pk_, inputs[i], n_in, aux ? aux[i] : nullptr, n_aux, r + 4 * (size_t)i, s + 4 * (size_t)i,
                        proofs_raw + 256 * (size_t)i, nullptr);
      if (rc) return rc;
    }
    return FB_OK;
  }
  std::vector<int> rcs(K, FB_OK);
  std::vector<std::string> errs(K);
  std::vector<std::thread> th;
  for (int t = 0; t < K; t++) {
    th.emplace_back([&, t]() {
      ProvingKey* p = t == 0 ? pk : pk->slots[t - 1];
      Ctx* c = t == 0 ? c0 : p->ctx;
      for (uint32_t i = t; i < count; i += K) {
        int rc = prove_impl(c, p, inputs[i], n_in, aux ? aux[i] : nullptr, n_aux, nullptr, r + 4 * (size_t)i,
                            s + 4 * (size_t)i, proofs_raw + 256 * (size_t)i, nullptr, nullptr);
        if (rc) { rcs[t] = rc; errs[t] = last_error_cstr(); return; }
      }
    });
  }
  for (auto& x : th) x.join();
  for (int t = 0; t < K; t++)
    if (rcs[t]) { set_error("%s", errs[t].c_str()); return rcs[t]; }
  return FB_OK;
} FB_ABI_CATCH_INT

// ---- streaming proves ------------------------------------------------
// The reference's prove() is synchronous: witness generation (prover.rs:69-76, host, single thread) and
// create_random_proof (prover.rs:78-80) alternate, so the prover idles while the next witness is made.
// A stream keeps up to `depth` submitted proofs queued or running on the key's batch slots; submit copies
// the witness and returns, so the caller generates witness k+1 while proof k runs, and collects proofs by
// ticket in any order.  Proof bytes are those of fb_prove on the same arguments.
struct ProveStream {
  struct Job {
    uint64_t ticket;
    int buf;  // index of the pinned witness buffer [inputs | aux] this proof reads
    uint64_t r[4], s[4];
  };
  struct Result {
    int rc = FB_OK;
    std::string err;
    uint8_t proof[256];
  };
  Ctx* c0 = nullptr;
  ProvingKey* pk = nullptr;
  std::mutex mu;
  std::condition_variable cv_job, cv_done, cv_space;
  std::deque<Job> queue;
  std::map<uint64_t, Result> done;
  std::set<uint64_t> outstanding;  // submitted and not collected yet
  uint64_t next_ticket = 1;
  size_t in_flight = 0, depth = 1;
  int waiters = 0;  // threads inside fb_stream_wait
  bool closing = false;
  std::vector<std::thread> workers;
  // one pinned witness buffer per proof that can be queued or running: submit is a plain memcpy (no page
  // faults) and the upload inside the prove is a real asynchronous DMA
  std::vector<uint64_t*> bufs;
  std::vector<int> free_bufs;
  // small keys: ONE worker takes whatever is queued (up to max_chunk proofs) and proves it as one batched chunk
  // (prove_batched: one set of launches, buckets keyed by proof) -- the batch size adapts to how fast the caller submits
  bool batched = false;
  uint32_t max_chunk = 64;

  void run_batched() {
    cudaSetDevice(c0->device);
    for (;;) {
      std::vector<Job> jobs;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_job.wait(lk, [&] { return closing || !queue.empty(); });
        if (queue.empty()) return;  // closing
        while (!queue.empty() && jobs.size() < max_chunk) {
          jobs.push_back(queue.front());
          queue.pop_front();
        }
      }
      const size_t n = jobs.size();
      std::vector<Result> res(n);
      if (n == 1) {
        const uint64_t* w = bufs[jobs[0].buf];
        res[0].rc = prove_impl(c0, pk, w, pk->n_in, pk->n_aux ? w + 4 * (size_t)pk->n_in : nullptr, pk->n_aux, nullptr,
                               jobs[0].r, jobs[0].s, res[0].proof, nullptr, nullptr);
        if (res[0].rc) res[0].err = last_error_cstr();
      } else {
        std::vector<const uint64_t*> ins(n), axs(n);
        std::vector<uint64_t> rr(4 * n), ss(4 * n);
        std::vector<uint8_t> proofs(256 * n);
        for (size_t i = 0; i < n; i++) {
          ins[i] = bufs[jobs[i].buf];
          axs[i] = ins[i] + 4 * (size_t)pk->n_in;
          memcpy(&rr[4 * i], jobs[i].r, 32);
          memcpy(&ss[4 * i], jobs[i].s, 32);
        }
        const int rc = prove_batched(c0, pk, (uint32_t)n, ins.data(), axs.data(), rr.data(), ss.data(), proofs.data());
        const std::string err = rc ? last_error_cstr() : "";
        for (size_t i = 0; i < n; i++) {
          res[i].rc = rc;
          res[i].err = err;
          memcpy(res[i].proof, &proofs[256 * i], 256);
        }
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        for (size_t i = 0; i < n; i++) {
          done.emplace(jobs[i].ticket, std::move(res[i]));
          free_bufs.push_back(jobs[i].buf);
          in_flight--;
        }
      }
      cv_done.notify_all();
      cv_space.notify_all();
    }
  }

  void run(int t) {
    ProvingKey* p = t == 0 ? pk : pk->slots[t - 1];
    Ctx* c = t == 0 ? c0 : p->ctx;
    cudaSetDevice(c0->device);
    for (;;) {
      Job job;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_job.wait(lk, [&] { return closing || !queue.empty(); });
        if (queue.empty()) return;  // closing
        job = queue.front();
        queue.pop_front();
      }
      Result res;
      const uint64_t* w = bufs[job.buf];
      res.rc = prove_impl(c, p, w, pk->n_in, pk->n_aux ? w + 4 * (size_t)pk->n_in : nullptr, pk->n_aux, nullptr, job.r,
                          job.s, res.proof, nullptr, nullptr);
      if (res.rc) res.err = last_error_cstr();
      {
        std::lock_guard<std::mutex> lk(mu);
        done.emplace(job.ticket, std::move(res));
        free_bufs.push_back(job.buf);
        in_flight--;
      }
      cv_done.notify_all();
      cv_space.notify_all();
    }
  }
};

int fb_stream_open(fb_ctx* ctx, fb_pk* pk_, int depth, fb_stream** out) try {
  ProvingKey* pk = reinterpret_cast<ProvingKey*>(pk_);
  Ctx* c0 = reinterpret_cast<Ctx*>(ctx);
  if (!pk || !c0 || !out) { set_error("fb_stream_open: null handle"); return FB_ERR_ARG; }
  if (pk->nshards != 1 || pk->dist_g) { set_error("fb_stream_open on a sharded key"); return FB_ERR_ARG; }
  int want = 8;
  if (const char* e = getenv("FB_BATCH_SLOTS")) want = std::max(1, std::min(32, atoi(e)));
  if (pk->m > (1u << 16) || g_serial) want = 1;  // a big prove fills the GPU on its own (fb_prove_batch)
  FB_CUDA(cudaSetDevice(c0->device));
  const char* mode = getenv("FB_BATCH_MODE");
  const bool batched = pk->m <= (1u << 16) && !g_serial && !(mode && !strcmp(mode, "slots"));
  uint32_t max_chunk = 64;
  if (const char* e = getenv("FB_BATCH_P")) max_chunk = (uint32_t)std::max(1, std::min(1024, atoi(e)));
  if (batched) want = 1;  // one worker, batched chunks
  while ((int)pk->slots.size() + 1 < want) {
    ProvingKey* sl = make_slot(pk);
    if (!sl) break;
    pk->slots.push_back(sl);
  }
  const int K = std::min<int>(want, (int)pk->slots.size() + 1);
  ProveStream* st = new ProveStream();
  st->c0 = c0;
  st->pk = pk;
  st->batched = batched;
  st->max_chunk = max_chunk;
  st->depth = (size_t)std::max(depth > 0 ? depth : (batched ? 2 * (int)max_chunk : 2 * K), 1);
  const size_t wbytes = std::max<size_t>(((size_t)pk->n_in + pk->n_aux) * sizeof(Fr), 32);
  for (size_t i = 0; i < st->depth; i++) {
    void* b = nullptr;
    if (cudaMallocHost(&b, wbytes) != cudaSuccess) {
      cudaGetLastError();
      for (uint64_t* q : st->bufs) cudaFreeHost(q);
      delete st;
      set_error("fb_stream_open: cannot pin %zu bytes per queued witness", wbytes);
      return FB_ERR_CUDA;
    }
    st->bufs.push_back(reinterpret_cast<uint64_t*>(b));
    st->free_bufs.push_back((int)i);
  }
  if (batched) st->workers.emplace_back([st] { st->run_batched(); });
  else for (int t = 0; t < K; t++) st->workers.emplace_back([st, t] { st->run(t); });
  *out = reinterpret_cast<fb_stream*>(st);
  return FB_OK;
} FB_ABI_CATCH_INT

int fb_stream_submit(fb_stream* st_, const uint64_t* inputs, uint32_t n_in, const uint64_t* aux, uint32_t n_aux,
                     const uint64_t r[4], const uint64_t s[4], uint64_t* ticket) try {
  ProveStream* st = reinterpret_cast<ProveStream*>(st_);
  if (!st || !inputs || (!aux && n_aux) || !r || !s || !ticket) { set_error("fb_stream_submit: bad argument"); return FB_ERR_ARG; }
  if (n_in != st->pk->n_in || n_aux != st->pk->n_aux) {
    set_error("witness has n_in=%u n_aux=%u, key expects %u / %u", n_in, n_aux, st->pk->n_in, st->pk->n_aux);
    return FB_ERR_ARG;
  }
  ProveStream::Job job;
  memcpy(job.r, r, 32);
  memcpy(job.s, s, 32);
  {
    std::unique_lock<std::mutex> lk(st->mu);
    st->cv_space.wait(lk, [&] { return st->in_flight < st->depth; });
    job.buf = st->free_bufs.back();  // in_flight < depth == number of buffers: one is free
    st->free_bufs.pop_back();
    job.ticket = *ticket = st->next_ticket++;
    st->in_flight++;
    st->outstanding.insert(job.ticket);
  }
  // the copy runs outside the lock: the buffer is owned by this job until its proof is done
  uint64_t* w = st->bufs[job.buf];
  memcpy(w, inputs, (size_t)n_in * 32);
  if (n_aux) memcpy(w + 4 * (size_t)n_in, aux, (size_t)n_aux * 32);
  {
    std::lock_guard<std::mutex> lk(st->mu);
    st->queue.push_back(job);
  }
  st->cv_job.notify_one();
  return FB_OK;
} FB_ABI_CATCH_INT

int fb_stream_wait(fb_stream* st_, uint64_t ticket, uint8_t proof_raw[256]) try {
  ProveStream* st = reinterpret_cast<ProveStream*>(st_);
  if (!st || !proof_raw) { set_error("fb_stream_wait: bad argument"); return FB_ERR_ARG; }
  std::unique_lock<std::mutex> lk(st->mu);
  if (!st->outstanding.count(ticket)) { set_error("fb_stream_wait: unknown or already collected ticket"); return FB_ERR_ARG; }
  st->waiters++;
  st->cv_done.wait(lk, [&] { return st->done.count(ticket) != 0; });
  st->waiters--;
  st->outstanding.erase(ticket);
  ProveStream::Result res = std::move(st->done[ticket]);
  st->done.erase(ticket);
  const bool closing = st->closing;
  lk.unlock();
  if (closing) st->cv_done.notify_all();
  if (res.rc) { set_error("%s", res.err.c_str()); return res.rc; }
  memcpy(proof_raw, res.proof, 256);
  return FB_OK;
} FB_ABI_CATCH_INT

void fb_stream_close(fb_stream* st_) try {
  ProveStream* st = reinterpret_cast<ProveStream*>(st_);
  if (!st) return;
  {
    std::lock_guard<std::mutex> lk(st->mu);
    st->closing = true;
    // proofs not started yet are dropped: publish an error result for each, so a thread blocked in
    // fb_stream_wait on one of them returns instead of waiting for ever
    for (const ProveStream::Job& job : st->queue) {
      ProveStream::Result res;
      res.rc = FB_ERR_ARG;
      res.err = "fb_stream_close: the stream was closed before this proof started";
      st->done.emplace(job.ticket, std::move(res));
      st->free_bufs.push_back(job.buf);
    }
    st->in_flight -= st->queue.size();
    st->queue.clear();
  }
  st->cv_job.notify_all();
  st->cv_done.notify_all();
  for (auto& t : st->workers) t.join();
  {  // let waiters that were woken up leave fb_stream_wait before the object goes away
    std::unique_lock<std::mutex> lk(st->mu);
    st->cv_done.wait(lk, [&] { return st->waiters == 0; });
  }
  for (uint64_t* b : st->bufs) cudaFreeHost(b);
  delete st;
} FB_ABI_CATCH_VOID

int fb_prove_device(fb_ctx* ctx, fb_pk* pk, const void* dev_w, const uint64_t r[4],
                    const uint64_t s[4], uint8_t proof_raw[256]) try {
  if (!dev_w || !r || !s || !proof_raw) { set_error("fb_prove_device: null buffer"); return FB_ERR_ARG; }
  ProvingKey* p = reinterpret_cast<ProvingKey*>(pk);
  if (p && p->nshards != 1) {  // collective, see fb_prove: every rank passes its own device copy of the witness
    if (!ctx) { set_error("fb_prove_device: null handle"); return FB_ERR_ARG; }
    return prove_collective(reinterpret_cast<Ctx*>(ctx), p, nullptr, 0, nullptr, 0, dev_w, r, s, proof_raw);
  }
  return prove_impl(reinterpret_cast<Ctx*>(ctx), p, nullptr, 0, nullptr, 0, dev_w, r, s, proof_raw,
                    nullptr, nullptr);
} FB_ABI_CATCH_INT

int fb_prove_partial(fb_ctx* ctx, fb_pk* pk, const uint64_t* inputs, uint32_t n_in,
                     const uint64_t* aux, uint32_t n_aux, uint8_t partial[640]) try {
  if (!inputs || (!aux && n_aux) || !partial) { set_error("fb_prove_partial: null buffer"); return FB_ERR_ARG; }
  return prove_impl(reinterpret_cast<Ctx*>(ctx), reinterpret_cast<ProvingKey*>(pk), inputs, n_in, aux,
                    n_aux, nullptr, nullptr, nullptr, nullptr, partial, nullptr);
} FB_ABI_CATCH_INT

int fb_prove_finish(const uint8_t* bellman_params, size_t len, const uint8_t* partials, int nparts,
                    const uint64_t r[4], const uint64_t s[4], uint8_t proof_raw[256]) try {
  if (!bellman_params || !partials || nparts < 1 || !r || !s || !proof_raw) { set_error("fb_prove_finish: bad argument"); return FB_ERR_ARG; }
  ParamsView v;
  int rc = parse_params(bellman_params, len < 580 ? len : 580, v);
  (void)rc;  // only the fixed-size verifying-key prefix is needed
  if (len < 576) { set_error("Parameters truncated in verifying key"); return FB_ERR_FORMAT; }
  VkPoints vk;
  if (host_decode_g1(bellman_params, vk.alpha_g1) || host_decode_g1(bellman_params + 64, vk.beta_g1) ||
      host_decode_g2(bellman_params + 128, vk.beta_g2) || host_decode_g1(bellman_params + 384, vk.delta_g1) ||
      host_decode_g2(bellman_params + 448, vk.delta_g2)) {
    set_error("invalid verifying-key point");
    return FB_ERR_FORMAT;
  }
  H1 sum[4] = {H1::inf(), H1::inf(), H1::inf(), H1::inf()};
  H2 sum2 = H2::inf();
  for (int i = 0; i < nparts; i++) {
    const uint8_t* p = partials + (size_t)i * 640;
    for (int j = 0; j < 4; j++) {
      G1Affine a;
      memcpy(&a, p + 128 * j, 64);
      sum[j] = add(sum[j], h1_from(a));
    }
    G2Affine q;
    memcpy(&q, p + 512, 128);
    sum2 = add(sum2, h2_from(q));
  }
  return assemble(&vk, sum[0], sum[1], sum[2], sum[3], sum2, r, s, proof_raw);
} FB_ABI_CATCH_INT

int fb_circuit_csr(const fb_circuit* c_, int m, const uint32_t** rowptr, const uint32_t** col,
                   const uint32_t** cidx, uint64_t* nnz, const uint64_t** coef_table, uint64_t* ncoef) try {
  const Circuit* c = reinterpret_cast<const Circuit*>(c_);
  if (!c || m < 0 || m > 2) return FB_ERR_ARG;
  if (rowptr) *rowptr = c->csr.rowptr[m].data();
  if (col) *col = c->csr.col[m].data();
  if (cidx) *cidx = c->csr.cidx[m].data();
  if (nnz) *nnz = c->csr.col[m].size();
  if (coef_table) *coef_table = reinterpret_cast<const uint64_t*>(c->csr.coef.data());
  if (ncoef) *ncoef = c->csr.coef.size();
  return FB_OK;
} FB_ABI_CATCH_INT

uint64_t fb_launch_count(void) { return fb::g_launches; }
void fb_set_serial(int on) { fb::g_serial = on != 0; }
void fb_set_msm_tables(int mode) { fb::g_msm_tables = mode < 0 ? -1 : (mode ? 1 : 0); }
void fb_set_prove_graph(int on) { fb::g_prove_graph.store(on ? 1 : 0); }
void fb_kernel_stats_enable(int on) { fb::kstat_enable(on != 0); }
void fb_kernel_stats_reset(void) { fb::kstat_reset(); }
int fb_kernel_stats(int which, uint64_t* launches, double* total_ms) try {
  if (which < 0 || which >= KSTAT_KINDS) return FB_ERR_ARG;
  unsigned long long n = 0;
  fb::kstat_collect(which, &n, total_ms);
  if (launches) *launches = n;
  return FB_OK;
} FB_ABI_CATCH_INT

int fb_prove_timings(const fb_pk* pk, float ms[6]) try {
  if (!pk || !ms) return FB_ERR_ARG;
  for (int i = 0; i < 6; i++) ms[i] = g_timing.ms[i];
  return FB_OK;
} FB_ABI_CATCH_INT

}  // extern "C"
