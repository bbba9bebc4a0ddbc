// Proving-key ingest: bellman Parameters bytes -> HBM-resident Montgomery affine bases,
// gate blob -> CSR.
//
// Follows (reference side of the boundary):
//   Parameters framing / read           fawkes-crypto/src/backend/bellman_groth16/mod.rs:150-175
//   gate stream: brotli -> borsh, 37 B per term (32 B canonical LE coeff, u8 tag, u32 LE idx)
//                                        fawkes-crypto/src/circuit/r1cs/cs.rs:184-223,248-250
//   Index tag 0 = Input, 1 = Aux         fawkes-crypto/src/circuit/r1cs/lc.rs:144-149
//   variables: inputs (ONE first) then aux   backend/bellman_groth16/mod.rs:61-102
// bellman's own point encoding (big-endian uncompressed, 0x40 flag = infinity) is restated
// from bellman_ce 0.3.5 / pairing_ce 0.18.1, absent from the repo.
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <unordered_map>

#include "internal.h"

namespace fb {

static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
}
const char* last_error_cstr() { return g_err.c_str(); }

std::atomic<unsigned long long> g_launches{0};
namespace {
struct KStat {
  std::atomic<bool> enabled{false};
  std::mutex mu;  // batch slots and stream workers launch MSMs concurrently
  std::vector<cudaEvent_t> pool;
  struct Rec { int kind; cudaEvent_t a, b; };
  std::vector<Rec> open_recs, recs;
  unsigned long long count[KSTAT_KINDS] = {};
  double ms[KSTAT_KINDS] = {};
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
} g_ks;
}  // namespace
void kstat_begin(int kind, cudaStream_t st) {
  if (!g_ks.enabled) return;
  std::lock_guard<std::mutex> lk(g_ks.mu);
  KStat::Rec r{kind, g_ks.get(), g_ks.get()};
  cudaEventRecord(r.a, st);
  g_ks.open_recs.push_back(r);
}
void kstat_end(int kind, cudaStream_t st) {
  if (!g_ks.enabled) return;
  std::lock_guard<std::mutex> lk(g_ks.mu);
  for (size_t i = g_ks.open_recs.size(); i-- > 0;) {
    if (g_ks.open_recs[i].kind == kind) {
      cudaEventRecord(g_ks.open_recs[i].b, st);
      g_ks.recs.push_back(g_ks.open_recs[i]);
      g_ks.open_recs.erase(g_ks.open_recs.begin() + i);
      return;
    }
  }
}
void kstat_enable(bool on) { g_ks.enabled = on; }
bool kstat_enabled() { return g_ks.enabled; }
void kstat_reset() {
  std::lock_guard<std::mutex> lk(g_ks.mu);
  for (auto& r : g_ks.recs) { g_ks.pool.push_back(r.a); g_ks.pool.push_back(r.b); }
  g_ks.recs.clear();
  for (int i = 0; i < KSTAT_KINDS; i++) { g_ks.count[i] = 0; g_ks.ms[i] = 0; }
}
void kstat_collect(int kind, unsigned long long* launches, double* ms) {
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_ks.mu);
  for (auto& r : g_ks.recs) {
    float t = 0;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { g_ks.count[r.kind]++; g_ks.ms[r.kind] += t; }
    g_ks.pool.push_back(r.a);
    g_ks.pool.push_back(r.b);
  }
  g_ks.recs.clear();
  if (launches) *launches = g_ks.count[kind];
  if (ms) *ms = g_ks.ms[kind];
}

// ---------------------------------------------------------------- brotli ---
// libbrotlidec.so.1 ships without headers in this image: bind the three symbols we need.
typedef struct BrotliDecoderStateStruct BrotliDecoderState;
typedef BrotliDecoderState* (*fn_create)(void*, void*, void*);
typedef int (*fn_stream)(BrotliDecoderState*, size_t*, const uint8_t**, size_t*, uint8_t**, size_t*);
typedef void (*fn_destroy)(BrotliDecoderState*);

namespace {
struct BrotliApi {
  void* lib = nullptr;
  fn_create create = nullptr;
  fn_stream stream = nullptr;
  fn_destroy destroy = nullptr;
  std::string err;
};
const BrotliApi& brotli_api() {
  static BrotliApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    api.lib = dlopen("libbrotlidec.so.1", RTLD_NOW);
    if (!api.lib) { api.err = std::string("cannot load libbrotlidec.so.1: ") + dlerror(); return; }
    api.create = (fn_create)dlsym(api.lib, "BrotliDecoderCreateInstance");
    api.stream = (fn_stream)dlsym(api.lib, "BrotliDecoderDecompressStream");
    api.destroy = (fn_destroy)dlsym(api.lib, "BrotliDecoderDestroyInstance");
    if (!api.create || !api.stream || !api.destroy) api.err = "libbrotlidec.so.1 lacks the streaming API";
  });
  return api;
}
}  // namespace

// max_out: the most the stream may legitimately expand to (callers derive it from num_gates and the
// variable counts); a stream that wants more is rejected instead of growing without bound.
int brotli_decode(const uint8_t* in, size_t len, std::vector<uint8_t>& out, size_t max_out) {
  const BrotliApi& api = brotli_api();
  if (!api.err.empty()) { set_error("%s", api.err.c_str()); return FB_ERR_FORMAT; }
  BrotliDecoderState* st = api.create(nullptr, nullptr, nullptr);
  if (!st) { set_error("BrotliDecoderCreateInstance failed"); return FB_ERR_FORMAT; }
  out.clear();
  out.resize(std::min<size_t>(std::max<size_t>(len * 4, 1 << 16), std::max<size_t>(max_out, 1)));
  size_t avail_in = len, produced = 0;
  const uint8_t* next_in = in;
  int rc = FB_OK;
  for (;;) {
    size_t avail_out = out.size() - produced;
    uint8_t* next_out = out.data() + produced;
    const int r = api.stream(st, &avail_in, &next_in, &avail_out, &next_out, nullptr);
    produced = out.size() - avail_out;
    if (r == 1) break;  // BROTLI_DECODER_RESULT_SUCCESS
    if (r == 3) {       // NEEDS_MORE_OUTPUT
      if (out.size() >= max_out) {
        set_error("gate stream expands beyond %zu bytes (more than the announced gate count can hold)", max_out);
        rc = FB_ERR_FORMAT;
        break;
      }
      out.resize(std::min<size_t>(out.size() * 2, max_out));
      continue;
    }
    // r == 0 (corrupt stream) or r == 2 (input ends inside the stream).  The reference's Decompressor surfaces
    // an io error there and the prover would run on a truncated circuit; at a key-load boundary that is an error.
    set_error(r == 2 ? "gate stream is truncated (brotli needs more input after %zu bytes)"
                     : "gate stream is not valid brotli (decoder error after %zu output bytes)",
              r == 2 ? len : produced);
    rc = FB_ERR_FORMAT;
    break;
  }
  api.destroy(st);
  out.resize(rc == FB_OK ? produced : 0);
  return rc;
}

// ------------------------------------------------------------- gate parse ---
struct FrKey {
  uint64_t w[4];
  bool operator==(const FrKey& o) const { return !memcmp(w, o.w, 32); }
};
struct FrKeyHash {
  size_t operator()(const FrKey& k) const {
    uint64_t h = k.w[0] * 0x9E3779B97F4A7C15ull;
    h ^= (k.w[1] + 0xBF58476D1CE4E5B9ull) * 0x94D049BB133111EBull;
    h ^= (k.w[2] << 7) ^ (k.w[3] >> 3) ^ (k.w[3] * 0xD6E8FEB86659FD93ull);
    return (size_t)h;
  }
};

int parse_gates_to_csr(const uint8_t* raw, size_t len, uint32_t n_in, uint32_t n_aux, HostCsr& out) {
  static const uint32_t kDedupCap = 1u << 20;
  std::unordered_map<FrKey, uint32_t, FrKeyHash> dict;
  for (int m = 0; m < 3; m++) {
    out.rowptr[m].clear();
    out.rowptr[m].push_back(0);
    out.col[m].clear();
    out.cidx[m].clear();
  }
  out.coef.clear();
  Fr minus_one = neg(Fr::one());
  size_t pos = 0;
  uint32_t gates = 0;
  for (;;) {
    // a gate is only accepted if all three parts parse (cs.rs:215-223)
    size_t save[3] = {out.col[0].size(), out.col[1].size(), out.col[2].size()};
    size_t save_coef = out.coef.size();
    bool ok = true;
    for (int m = 0; m < 3 && ok; m++) {
      if (pos + 4 > len) { ok = false; break; }
      uint32_t cnt;
      memcpy(&cnt, raw + pos, 4);
      pos += 4;
      if ((size_t)cnt * 37 > len - pos) { ok = false; break; }
      for (uint32_t t = 0; t < cnt; t++) {
        Fr c;
        memcpy(c.v, raw + pos, 32);
        uint8_t tag = raw[pos + 32];
        uint32_t idx;
        memcpy(&idx, raw + pos + 33, 4);
        pos += 37;
        if (geq_mod<FrCfg>(c.v) || tag > 1) { ok = false; break; }  // "Wrong raw integer" / enum overflow
        if ((tag == 0 && idx >= n_in) || (tag == 1 && idx >= n_aux)) {
          set_error("gate %u references %s variable %u out of range", gates, tag ? "aux" : "input", idx);
          return FB_ERR_FORMAT;
        }
        Fr cm = to_mont(c);
        uint32_t ci;
        if (cm == Fr::one()) ci = 0;
        else if (cm == minus_one) ci = 1;
        else {
          FrKey key;
          memcpy(key.w, cm.v, 32);
          auto it = dict.find(key);
          if (it != dict.end()) ci = it->second;
          else {
            ci = (uint32_t)out.coef.size() + 2;
            out.coef.push_back(cm);
            if (dict.size() < kDedupCap) dict.emplace(key, ci);
          }
        }
        out.col[m].push_back(tag == 0 ? idx : n_in + idx);
        out.cidx[m].push_back(ci);
      }
    }
    if (!ok) {
      for (int m = 0; m < 3; m++) { out.col[m].resize(save[m]); out.cidx[m].resize(save[m]); }
      (void)save_coef;  // coefficients appended by a rejected gate stay unused
      break;
    }
    for (int m = 0; m < 3; m++) out.rowptr[m].push_back((uint32_t)out.col[m].size());
    gates++;
  }
  out.n_gates = gates;
  return FB_OK;
}

int upload_csr(const HostCsr& h, DevCsr& d, cudaStream_t st) {
  d.n_gates = h.n_gates;
  for (int m = 0; m < 3; m++) {
    d.nnz[m] = h.col[m].size();
    FB_CUDA(cudaMalloc(&d.rowptr[m], h.rowptr[m].size() * 4));
    FB_CUDA(cudaMalloc(&d.col[m], std::max<size_t>(h.col[m].size(), 1) * 4));
    FB_CUDA(cudaMalloc(&d.cidx[m], std::max<size_t>(h.cidx[m].size(), 1) * 4));
    FB_CUDA(cudaMemcpyAsync(d.rowptr[m], h.rowptr[m].data(), h.rowptr[m].size() * 4, cudaMemcpyHostToDevice, st));
    FB_CUDA(cudaMemcpyAsync(d.col[m], h.col[m].data(), h.col[m].size() * 4, cudaMemcpyHostToDevice, st));
    FB_CUDA(cudaMemcpyAsync(d.cidx[m], h.cidx[m].data(), h.cidx[m].size() * 4, cudaMemcpyHostToDevice, st));
  }
  FB_CUDA(cudaMalloc(&d.coef, std::max<size_t>(h.coef.size(), 1) * sizeof(Fr)));
  FB_CUDA(cudaMemcpyAsync(d.coef, h.coef.data(), h.coef.size() * sizeof(Fr), cudaMemcpyHostToDevice, st));
  FB_CUDA(cudaStreamSynchronize(st));
  return FB_OK;
}

void free_csr(DevCsr& d) {
  for (int m = 0; m < 3; m++) {
    cudaFree(d.rowptr[m]); cudaFree(d.col[m]); cudaFree(d.cidx[m]);
    d.rowptr[m] = d.col[m] = d.cidx[m] = nullptr;
  }
  cudaFree(d.coef);
  d.coef = nullptr;
}

// ----------------------------------------------------- point (de)coding ---
// 32 big-endian bytes -> canonical limbs
FB_HD void be_to_limbs(const uint8_t* be, uint32_t* v) {
  for (int i = 0; i < 8; i++) {
    const uint8_t* p = be + 32 - 4 * (i + 1);
    v[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
  }
}
FB_HD void limbs_to_be(const uint32_t* v, uint8_t* be) {
  for (int i = 0; i < 8; i++) {
    uint8_t* p = be + 32 - 4 * (i + 1);
    p[0] = v[i] >> 24; p[1] = v[i] >> 16; p[2] = v[i] >> 8; p[3] = v[i];
  }
}

// Decoder of pairing_ce's `Uncompressed` encoding as bellman's Parameters::read drives it (bellman_ce 0.3.5
// groth16/mod.rs `read_g1/read_g2` closures, restated -- the crate is not vendored):
//   unchecked (`into_affine_unchecked`): bit 7 (compression) must be clear; bit 6 = infinity, and then every other
//     bit of the encoding must be zero; coordinates must be < p;
//   checked   (`into_affine`): additionally on the curve and in the r-torsion subgroup (G1 has cofactor 1, so the
//     subgroup test only costs anything for G2: [r]P == O);
//   disallow_points_at_infinity: the point at infinity is an error.
// flags: FB_LOAD_CHECKED | FB_LOAD_NO_INFINITY.
// returns 0 ok, 1 = coordinate >= p, 2 = not on curve, 3 = bad flag, 4 = dirty infinity encoding,
// 5 = not in the r-torsion subgroup, 6 = point at infinity where none is allowed
template <int SZ>
FB_HD bool inf_encoding_clean(const uint8_t* be) {
  if (be[0] != 0x40) return false;
  for (int i = 1; i < SZ; i++)
    if (be[i]) return false;
  return true;
}
template <class F>
FB_HD_COLD bool in_r_torsion(const Affine<F>& p) {
  uint32_t r[8];
  for (int i = 0; i < 8; i++) r[i] = FrCfg::mod(i);
  return scalar_mul(XYZZ<F>::from_affine(p), r).is_inf();
}
FB_HD int decode_g1(const uint8_t* be, G1Affine& out, int flags) {
  if (be[0] & 0x80) return 3;
  if (be[0] & 0x40) {
    if (!inf_encoding_clean<64>(be)) return 4;
    if (flags & FB_LOAD_NO_INFINITY) return 6;
    out = G1Affine::inf();
    return 0;
  }
  Fq x, y;
  be_to_limbs(be, x.v);
  be_to_limbs(be + 32, y.v);
  if (geq_mod<FqCfg>(x.v) || geq_mod<FqCfg>(y.v)) return 1;
  out.x = to_mont(x);
  out.y = to_mont(y);
  // (0, 0) is the in-memory form of infinity (group.rs:55) and is not on the curve: never a valid encoding
  if ((flags & FB_LOAD_CHECKED) && !on_curve(out, g1_b())) return 2;
  if (out.is_inf()) return 2;
  return 0;
}
FB_HD int decode_g2(const uint8_t* be, G2Affine& out, int flags) {
  if (be[0] & 0x80) return 3;
  if (be[0] & 0x40) {
    if (!inf_encoding_clean<128>(be)) return 4;
    if (flags & FB_LOAD_NO_INFINITY) return 6;
    out = G2Affine::inf();
    return 0;
  }
  Fq c[4];
  for (int i = 0; i < 4; i++) {
    be_to_limbs(be + 32 * i, c[i].v);
    if (geq_mod<FqCfg>(c[i].v)) return 1;
    c[i] = to_mont(c[i]);
  }
  out.x = {c[1], c[0]};  // bytes are x.c1 | x.c0 | y.c1 | y.c0
  out.y = {c[3], c[2]};
  if (out.is_inf()) return 2;
  if (flags & FB_LOAD_CHECKED) {
    if (!on_curve(out, g2_b())) return 2;
    if (!in_r_torsion(out)) return 5;
  }
  return 0;
}

__global__ void k_decode_g1(const uint8_t* __restrict__ be, uint64_t n, G1Affine* __restrict__ out,
                            int flags, int* __restrict__ err) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    G1Affine p;
    int e = decode_g1(be + i * 64, p, flags);
    if (e) atomicMax(err, e);
    else out[i] = p;
  }
}
__global__ void k_decode_g2(const uint8_t* __restrict__ be, uint64_t n, G2Affine* __restrict__ out,
                            int flags, int* __restrict__ err) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    G2Affine p;
    int e = decode_g2(be + i * 128, p, flags);
    if (e) atomicMax(err, e);
    else out[i] = p;
  }
}

static const char* point_error_name(int e) {
  switch (e) {
    case 1: return "coordinate not in field";
    case 2: return "not on curve";
    case 3: return "unexpected compression flag";
    case 4: return "unexpected information in the encoding of the point at infinity";
    case 5: return "not in the r-torsion subgroup";
    case 6: return "point at infinity";
    default: return "invalid";
  }
}

namespace {
struct DevBuf {  // cudaFree on every exit path
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
};
}  // namespace

template <class P, int SZ, class K>
static int decode_points(const uint8_t* host_be, uint64_t n, P* dev_out, int flags,
                         cudaStream_t st, K kernel) {
  if (n == 0) return FB_OK;
  // stage through a bounded device buffer so a 16 GiB key does not need 16 GiB of staging
  const uint64_t chunk = std::min<uint64_t>(n, 1ull << 22);
  DevBuf stage_b, derr_b;
  FB_CUDA(cudaMalloc(&stage_b.p, chunk * SZ));
  FB_CUDA(cudaMalloc(&derr_b.p, 4));
  uint8_t* stage = reinterpret_cast<uint8_t*>(stage_b.p);
  int* derr = reinterpret_cast<int*>(derr_b.p);
  FB_CUDA(cudaMemsetAsync(derr, 0, 4, st));
  for (uint64_t off = 0; off < n; off += chunk) {
    uint64_t cnt = std::min(chunk, n - off);
    FB_CUDA(cudaMemcpyAsync(stage, host_be + off * SZ, cnt * SZ, cudaMemcpyHostToDevice, st));
    unsigned blocks = (unsigned)std::min<uint64_t>((cnt + 127) / 128, 148 * 16);
    kernel<<<blocks, 128, 0, st>>>(stage, cnt, dev_out + off, flags, derr);
    FB_CUDA(cudaStreamSynchronize(st));
  }
  int herr = 0;
  FB_CUDA(cudaMemcpyAsync(&herr, derr, 4, cudaMemcpyDeviceToHost, st));
  FB_CUDA(cudaStreamSynchronize(st));
  if (herr) {
    set_error("invalid point in Parameters (%s)", point_error_name(herr));
    return FB_ERR_FORMAT;
  }
  return FB_OK;
}

int decode_g1_be(const uint8_t* host_be, uint64_t n, G1Affine* dev_out, int flags, cudaStream_t st) {
  return decode_points<G1Affine, 64>(host_be, n, dev_out, flags, st, k_decode_g1);
}
int decode_g2_be(const uint8_t* host_be, uint64_t n, G2Affine* dev_out, int flags, cudaStream_t st) {
  return decode_points<G2Affine, 128>(host_be, n, dev_out, flags, st, k_decode_g2);
}
// verifying-key points are always read checked (bellman VerifyingKey::read uses into_affine)
int host_decode_g1(const uint8_t* be, G1Affine& out, int extra_flags) { return decode_g1(be, out, FB_LOAD_CHECKED | extra_flags); }
int host_decode_g2(const uint8_t* be, G2Affine& out, int extra_flags) { return decode_g2(be, out, FB_LOAD_CHECKED | extra_flags); }
const char* point_error(int e) { return point_error_name(e); }

void host_encode_g1(const G1Affine& p, uint8_t* be) {
  memset(be, 0, 64);
  if (p.is_inf()) { be[0] = 0x40; return; }
  Fq x = from_mont(p.x), y = from_mont(p.y);
  limbs_to_be(x.v, be);
  limbs_to_be(y.v, be + 32);
}
void host_encode_g2(const G2Affine& p, uint8_t* be) {
  memset(be, 0, 128);
  if (p.is_inf()) { be[0] = 0x40; return; }
  Fq c[4] = {from_mont(p.x.c1), from_mont(p.x.c0), from_mont(p.y.c1), from_mont(p.y.c0)};
  for (int i = 0; i < 4; i++) limbs_to_be(c[i].v, be + 32 * i);
}

}  // namespace fb
