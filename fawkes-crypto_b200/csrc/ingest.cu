// Gate-blob ingest on the GPU: borsh gate stream -> CSR + coefficient dictionary.
//
// The reference re-parses the stream on every prove: brotli -> borsh, 37 bytes per term (32 B canonical
// little-endian coefficient, u8 Index tag, u32 index), one canonical->Montgomery multiply per coefficient
//   fawkes-crypto/src/circuit/r1cs/cs.rs:184-223 (GateStreamedIterator), :248-250 (get_gate_iterator)
//   fawkes-crypto/src/circuit/r1cs/lc.rs:144-149 (Index: 0 = Input, 1 = Aux)
//   ff-uint_derive/src/lib.rs:687-702 (borsh of a field element = canonical LE, "Wrong raw integer" when >= p)
// Here it is parsed once per key.  What stays on the host is only what is sequential by construction: brotli,
// and the walk over the length prefixes that finds where every LC starts (one u32 read per LC).  The per-term
// work -- unaligned 37-byte reads, range checks, the Montgomery multiply, the +1 / -1 / dictionary coding of
// the coefficient -- runs as kernels over all terms at once:
//
//   k_ingest_terms    one thread per LC: parse, validate, write col[], classify the coefficient
//   scan              stream-order rank of every dictionary ("other") coefficient
//   k_ingest_collect  Montgomery values of the dictionary candidates, in stream order
//   k_ingest_hash     open-addressing table keyed by the 256-bit value; every class keeps its EARLIEST member
//   scan              dense dictionary index of the class representatives (first occurrence order)
//   k_ingest_assign   cidx[] of every term, coef[] of every representative
//
// The result is the same HostCsr the host parser (pk.cu: parse_gates_to_csr) builds -- byte for byte as long
// as that parser's dictionary cap (2^20 distinct values) is not reached, because both number the distinct
// coefficients by first appearance in the stream (tests/test_gpu_prove.py::test_gpu_ingest_*).
#include <algorithm>
#include <chrono>
#include <cstring>
#include <thread>

#include "internal.h"

namespace fb {

namespace {

constexpr unsigned long long kNoErr = ~0ull;
constexpr uint32_t kEmpty = 0xffffffffu;

struct IngestArgs {
  const uint32_t* raw;         // blob, 4-byte aligned, padded with >= 8 readable bytes
  const uint64_t* gstart;      // byte offset of every gate
  const uint32_t* rp[3];       // row pointers of A, B, C (n_gates + 1 each)
  uint32_t* col[3];
  uint32_t* cidx[3];
  uint32_t n_gates, n_in, n_aux;
};

struct Term {
  Fr c;          // canonical
  uint32_t tag, idx;
};

// 37 unaligned bytes at byte offset `off`
__device__ __forceinline__ Term read_term(const uint32_t* __restrict__ raw, uint64_t off) {
  const uint32_t* p = raw + (off >> 2);
  const uint32_t sh = (uint32_t)(off & 3) * 8;
  uint32_t a[11];
#pragma unroll
  for (int i = 0; i < 11; i++) a[i] = __ldg(p + i);
  uint32_t w[10];
#pragma unroll
  for (int i = 0; i < 10; i++) w[i] = __funnelshift_r(a[i], a[i + 1], sh);
  Term t;
#pragma unroll
  for (int i = 0; i < 8; i++) t.c.v[i] = w[i];
  t.tag = w[8] & 0xffu;
  t.idx = (w[8] >> 8) | (w[9] << 24);
  return t;
}

// where LC (g, m) lives: byte offset of its first term, its first slot in col[m] / cidx[m], its first
// position in stream order (all terms of A, B, C interleaved gate by gate) and its length
struct LcPos {
  uint64_t off;
  uint32_t t0, s0, cnt;
};
__device__ __forceinline__ LcPos lc_pos(const IngestArgs& a, uint32_t g, int m) {
  const uint32_t b0 = a.rp[0][g], b1 = a.rp[1][g], b2 = a.rp[2][g];
  const uint32_t c0 = a.rp[0][g + 1] - b0, c1 = a.rp[1][g + 1] - b1, c2 = a.rp[2][g + 1] - b2;
  LcPos p;
  p.off = a.gstart[g] + 4;
  p.s0 = b0 + b1 + b2;
  p.t0 = b0;
  p.cnt = c0;
  if (m >= 1) { p.off += 37ull * c0 + 4; p.s0 += c0; p.t0 = b1; p.cnt = c1; }
  if (m == 2) { p.off += 37ull * c1 + 4; p.s0 += c1; p.t0 = b2; p.cnt = c2; }
  return p;
}

// err[0]: earliest byte offset of a term that does not parse (coefficient >= r, tag > 1: the reference's
// iterator ends there); err[1]: earliest offset of a term whose variable index is out of range
__global__ void __launch_bounds__(256)
k_ingest_terms(IngestArgs a, uint32_t* __restrict__ other, unsigned long long* __restrict__ err) {
  const uint64_t lc = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lc >= 3ull * a.n_gates) return;
  const uint32_t g = (uint32_t)(lc / 3);
  const int m = (int)(lc % 3);
  const LcPos p = lc_pos(a, g, m);
  const Fr one = Fr::one(), minus_one = neg(Fr::one());
  for (uint32_t k = 0; k < p.cnt; k++) {
    const uint64_t off = p.off + 37ull * k;
    const Term t = read_term(a.raw, off);
    uint32_t ci = 0, col = 0, oth = 0;
    if (geq_mod<FrCfg>(t.c.v) || t.tag > 1) {
      atomicMin(&err[0], (unsigned long long)off);
    } else if ((t.tag == 0 && t.idx >= a.n_in) || (t.tag == 1 && t.idx >= a.n_aux)) {
      atomicMin(&err[1], (unsigned long long)off);
    } else {
      col = t.tag == 0 ? t.idx : a.n_in + t.idx;
      const Fr cm = to_mont(t.c);
      if (cm == one) ci = 0;
      else if (cm == minus_one) ci = 1;
      else oth = 1;
    }
    a.col[m][p.t0 + k] = col;
    a.cidx[m][p.t0 + k] = ci;
    other[p.s0 + k] = oth;
  }
}

// candidates in stream order: vals[o] = Montgomery value, where[o] = slot in cidx (matrix in the top 2 bits)
__global__ void __launch_bounds__(256)
k_ingest_collect(IngestArgs a, const uint32_t* __restrict__ other, const uint32_t* __restrict__ opos,
                 Fr* __restrict__ vals, uint32_t* __restrict__ where) {
  const uint64_t lc = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lc >= 3ull * a.n_gates) return;
  const uint32_t g = (uint32_t)(lc / 3);
  const int m = (int)(lc % 3);
  const LcPos p = lc_pos(a, g, m);
  for (uint32_t k = 0; k < p.cnt; k++) {
    if (!other[p.s0 + k]) continue;
    const uint32_t o = opos[p.s0 + k];
    vals[o] = to_mont(read_term(a.raw, p.off + 37ull * k).c);
    where[o] = (p.t0 + k) | ((uint32_t)m << 30);
  }
}

__device__ __forceinline__ uint32_t hash_fr(const Fr& x) {
  uint64_t h = 0x9E3779B97F4A7C15ull;
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    const uint64_t w = x.v[i] | ((uint64_t)x.v[i + 1] << 32);
    h = (h ^ w) * 0xBF58476D1CE4E5B9ull;
    h ^= h >> 29;
  }
  h *= 0x94D049BB133111EBull;
  return (uint32_t)(h >> 32);
}

// table[i] always holds a member of ONE equivalence class (or kEmpty), and only ever moves to an earlier
// member of the same class, so comparing against whatever is read stays valid under concurrency.
__global__ void __launch_bounds__(256)
k_ingest_hash(const Fr* __restrict__ vals, uint32_t n, uint32_t* __restrict__ table, uint32_t mask,
              uint32_t* __restrict__ slot_of) {
  const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n) return;
  const Fr v = vals[o];
  uint32_t i = hash_fr(v) & mask;
  for (;;) {
    uint32_t cur = atomicCAS(&table[i], kEmpty, o);
    if (cur == kEmpty) break;                       // claimed a fresh slot
    if (vals[cur] == v) { atomicMin(&table[i], o); break; }
    i = (i + 1) & mask;
  }
  slot_of[o] = i;
}

__global__ void __launch_bounds__(256)
k_ingest_isrep(const uint32_t* __restrict__ table, const uint32_t* __restrict__ slot_of, uint32_t n,
               uint32_t* __restrict__ isrep) {
  const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o < n) isrep[o] = table[slot_of[o]] == o ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
k_ingest_assign(const Fr* __restrict__ vals, const uint32_t* __restrict__ table, const uint32_t* __restrict__ slot_of,
                const uint32_t* __restrict__ dpos, const uint32_t* __restrict__ where, uint32_t n,
                uint32_t* cidx0, uint32_t* cidx1, uint32_t* cidx2, Fr* __restrict__ coef) {
  const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n) return;
  const uint32_t rep = table[slot_of[o]];
  const uint32_t d = dpos[rep];
  const uint32_t w = where[o];
  uint32_t* cidx = (w >> 30) == 0 ? cidx0 : (w >> 30) == 1 ? cidx1 : cidx2;
  cidx[w & 0x3fffffffu] = 2 + d;
  if (rep == o) coef[d] = vals[o];
}

// ---- exclusive scan of u32 (three kernels, any length below 2^32) ----
__global__ void __launch_bounds__(1024)
k_iscan_block(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t* __restrict__ sums, uint32_t n) {
  __shared__ uint32_t wsum[32];
  const uint64_t i = (uint64_t)blockIdx.x * 1024 + threadIdx.x;
  const uint32_t v = i < n ? in[i] : 0;
  uint32_t x = v;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) wsum[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = wsum[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += y;
    }
    wsum[lane] = w;
  }
  __syncthreads();
  const uint32_t incl = x + (wid ? wsum[wid - 1] : 0);
  if (i < n) out[i] = incl - v;
  if (threadIdx.x == 1023) sums[blockIdx.x] = incl;
}
__global__ void __launch_bounds__(1024) k_iscan_sums(uint32_t* sums, uint32_t nb) {  // one CTA, serial over chunks
  __shared__ uint32_t carry;
  __shared__ uint32_t wsum[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nb; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < nb ? sums[i] : 0;
    uint32_t x = v;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) wsum[wid] = x;
    __syncthreads();
    if (wid == 0) {
      uint32_t w = wsum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += y;
      }
      wsum[lane] = w;
    }
    __syncthreads();
    const uint32_t incl = x + (wid ? wsum[wid - 1] : 0) + carry;
    if (i < nb) sums[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[nb] = carry;  // grand total
}
__global__ void __launch_bounds__(1024)
k_iscan_add(uint32_t* __restrict__ out, const uint32_t* __restrict__ sums, uint32_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * 1024 + threadIdx.x;
  if (i < n) out[i] += sums[blockIdx.x];
}

struct DevBuf {  // frees on scope exit
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 16)); }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// out = exclusive scan of in (n entries); *total = sum.  sums: scratch of n/1024 + 2 entries.
int scan_u32(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* sums, uint32_t* total, cudaStream_t st) {
  if (n == 0) { *total = 0; return FB_OK; }
  const uint32_t nb = (uint32_t)(((uint64_t)n + 1023) / 1024);
  k_iscan_block<<<nb, 1024, 0, st>>>(in, out, sums, n);
  k_iscan_sums<<<1, 1024, 0, st>>>(sums, nb);
  k_iscan_add<<<nb, 1024, 0, st>>>(out, sums, n);
  FB_CUDA(cudaMemcpyAsync(total, sums + nb, 4, cudaMemcpyDeviceToHost, st));
  FB_CUDA(cudaStreamSynchronize(st));
  return FB_OK;
}

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// Walk the length prefixes: gate starts and the three row-pointer arrays.  Stops at the first gate whose
// framing does not fit the buffer (cs.rs:215-223: a gate exists only if all three LCs deserialize).
static uint32_t frame_gates(const uint8_t* raw, size_t len, uint32_t max_gates, std::vector<uint64_t>& gstart,
                            std::vector<uint32_t> rp[3], bool* overflow) {
  size_t pos = 0;
  uint32_t gates = 0;
  uint64_t nnz[3] = {0, 0, 0};
  for (int m = 0; m < 3; m++) { rp[m].clear(); rp[m].push_back(0); }
  gstart.clear();
  *overflow = false;
  while (gates < max_gates) {
    const size_t g0 = pos;
    uint32_t cnt[3];
    bool ok = true;
    for (int m = 0; m < 3; m++) {
      if (pos + 4 > len) { ok = false; break; }
      memcpy(&cnt[m], raw + pos, 4);
      pos += 4;
      if ((size_t)cnt[m] * 37 > len - pos) { ok = false; break; }
      pos += (size_t)cnt[m] * 37;
    }
    if (!ok) break;
    for (int m = 0; m < 3; m++) {
      nnz[m] += cnt[m];
      if (nnz[m] >= (1ull << 30)) { *overflow = true; return gates; }
      rp[m].push_back((uint32_t)nnz[m]);
    }
    gstart.push_back(g0);
    gates++;
  }
  return gates;
}

// times_ms (optional, 6 entries): framing walk on the host and, concurrently, blob upload; kernels; copy back, [4] unused here
// (brotli, filled by the caller), cudaMalloc time
int parse_gates_device(Ctx* ctx, const uint8_t* raw, size_t len, uint32_t n_in, uint32_t n_aux, HostCsr& out,
                       float* times_ms) {
  if (!ctx) { set_error("parse_gates_device: no context"); return FB_ERR_ARG; }
  FB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  double t0 = now_s();
  // the length prefixes are walked on a helper thread (plain CPU work) while this one uploads the blob.
  // (The other way round -- CUDA calls on a short-lived helper thread -- made the first cudaMalloc after
  // the thread's exit take ~140 ms in two of three runs.)
  DevBuf d_raw;
  const size_t raw_words = (len + 3) / 4 + 3;  // read_term touches up to 11 words from an aligned base
  float t_up = 0, t_kern = 0, t_down = 0;
  double t_alloc = 0, ta = 0, t_frame = 0;
  std::vector<uint64_t> gstart;
  bool overflow = false;
  uint32_t n_gates = 0;
  std::thread walker([&] {
    n_gates = frame_gates(raw, len, 0xffffffffu, gstart, out.rowptr, &overflow);
    t_frame = now_s() - t0;
  });
  cudaError_t up_err = cudaSuccess;
  if (len) {
    up_err = d_raw.alloc(raw_words * 4);
    if (up_err == cudaSuccess)
      up_err = cudaMemsetAsync(d_raw.as<uint8_t>() + (len / 4) * 4, 0, raw_words * 4 - (len / 4) * 4, st);
    if (up_err == cudaSuccess) up_err = cudaMemcpyAsync(d_raw.p, raw, len, cudaMemcpyHostToDevice, st);
    if (up_err == cudaSuccess) up_err = cudaStreamSynchronize(st);
    t_up = (float)((now_s() - t0) * 1e3);
  }
  walker.join();
  FB_CUDA(up_err);
  if (overflow) { set_error("gate stream holds 2^30 or more terms in one matrix"); return FB_ERR_FORMAT; }

  for (int attempt = 0; attempt < 2 && n_gates; attempt++) {
    t0 = now_s();
    const uint64_t nnz[3] = {out.rowptr[0][n_gates], out.rowptr[1][n_gates], out.rowptr[2][n_gates]};
    const uint64_t total64 = nnz[0] + nnz[1] + nnz[2];
    if (total64 >= 0xffffffffull) { set_error("gate stream holds 2^32 or more terms"); return FB_ERR_FORMAT; }
    const uint32_t total = (uint32_t)total64;
    DevBuf d_gstart, d_rp[3], d_col[3], d_cidx[3], d_other, d_opos, d_sums, d_err;
    ta = now_s();
    FB_CUDA(d_gstart.alloc((size_t)n_gates * 8));
    for (int m = 0; m < 3; m++) {
      FB_CUDA(d_rp[m].alloc(((size_t)n_gates + 1) * 4));
      FB_CUDA(d_col[m].alloc(nnz[m] * 4));
      FB_CUDA(d_cidx[m].alloc(nnz[m] * 4));
    }
    FB_CUDA(d_other.alloc((size_t)total * 4));
    FB_CUDA(d_opos.alloc((size_t)total * 4));
    FB_CUDA(d_sums.alloc(((size_t)total / 1024 + 4) * 4));
    FB_CUDA(d_err.alloc(16));
    t_alloc += now_s() - ta;
    FB_CUDA(cudaMemcpyAsync(d_gstart.p, gstart.data(), (size_t)n_gates * 8, cudaMemcpyHostToDevice, st));
    IngestArgs a;
    a.raw = d_raw.as<uint32_t>();
    a.gstart = d_gstart.as<uint64_t>();
    a.n_gates = n_gates;
    a.n_in = n_in;
    a.n_aux = n_aux;
    for (int m = 0; m < 3; m++) {
      FB_CUDA(cudaMemcpyAsync(d_rp[m].p, out.rowptr[m].data(), ((size_t)n_gates + 1) * 4, cudaMemcpyHostToDevice, st));
      a.rp[m] = d_rp[m].as<uint32_t>();
      a.col[m] = d_col[m].as<uint32_t>();
      a.cidx[m] = d_cidx[m].as<uint32_t>();
    }
    const unsigned long long no_err[2] = {kNoErr, kNoErr};
    FB_CUDA(cudaMemcpyAsync(d_err.p, no_err, 16, cudaMemcpyHostToDevice, st));
    const unsigned lc_blocks = (unsigned)((3ull * n_gates + 255) / 256);
    k_ingest_terms<<<lc_blocks, 256, 0, st>>>(a, d_other.as<uint32_t>(), d_err.as<unsigned long long>());
    unsigned long long err[2];
    FB_CUDA(cudaMemcpyAsync(err, d_err.p, 16, cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaStreamSynchronize(st));
    FB_CUDA(cudaGetLastError());
    if (err[1] < err[0]) {  // the first bad thing in stream order is an index out of range: a format error
      const size_t off = (size_t)err[1];
      const uint32_t g = (uint32_t)(std::upper_bound(gstart.begin(), gstart.begin() + n_gates, (uint64_t)off) - gstart.begin()) - 1;
      uint32_t idx;
      memcpy(&idx, raw + off + 33, 4);
      set_error("gate %u references %s variable %u out of range", g, raw[off + 32] ? "aux" : "input", idx);
      return FB_ERR_FORMAT;
    }
    if (err[0] != kNoErr) {
      // the stream ends in front of the gate that holds the first unparsable term: redo on the prefix
      const uint32_t g = (uint32_t)(std::upper_bound(gstart.begin(), gstart.begin() + n_gates, (uint64_t)err[0]) - gstart.begin()) - 1;
      n_gates = g;
      for (int m = 0; m < 3; m++) out.rowptr[m].resize((size_t)n_gates + 1);
      continue;
    }
    // dictionary of the coefficients that are neither +1 nor -1
    uint32_t n_other = 0, n_coef = 0;
    int rc = scan_u32(d_other.as<uint32_t>(), d_opos.as<uint32_t>(), total, d_sums.as<uint32_t>(), &n_other, st);
    if (rc) return rc;
    DevBuf d_vals, d_where, d_table, d_slot, d_isrep, d_dpos, d_coef;
    if (n_other) {
      uint32_t tsize = 1024;
      while (tsize < 2ull * n_other && tsize < (1u << 31)) tsize <<= 1;
      ta = now_s();
      FB_CUDA(d_vals.alloc((size_t)n_other * sizeof(Fr)));
      FB_CUDA(d_where.alloc((size_t)n_other * 4));
      FB_CUDA(d_table.alloc((size_t)tsize * 4));
      FB_CUDA(d_slot.alloc((size_t)n_other * 4));
      FB_CUDA(d_isrep.alloc((size_t)n_other * 4));
      FB_CUDA(d_dpos.alloc((size_t)n_other * 4));
      t_alloc += now_s() - ta;
      FB_CUDA(cudaMemsetAsync(d_table.p, 0xff, (size_t)tsize * 4, st));
      k_ingest_collect<<<lc_blocks, 256, 0, st>>>(a, d_other.as<uint32_t>(), d_opos.as<uint32_t>(), d_vals.as<Fr>(),
                                                  d_where.as<uint32_t>());
      const unsigned ob = (n_other + 255) / 256;
      k_ingest_hash<<<ob, 256, 0, st>>>(d_vals.as<Fr>(), n_other, d_table.as<uint32_t>(), tsize - 1, d_slot.as<uint32_t>());
      k_ingest_isrep<<<ob, 256, 0, st>>>(d_table.as<uint32_t>(), d_slot.as<uint32_t>(), n_other, d_isrep.as<uint32_t>());
      rc = scan_u32(d_isrep.as<uint32_t>(), d_dpos.as<uint32_t>(), n_other, d_sums.as<uint32_t>(), &n_coef, st);
      if (rc) return rc;
      ta = now_s();
      FB_CUDA(d_coef.alloc((size_t)n_coef * sizeof(Fr)));
      t_alloc += now_s() - ta;
      k_ingest_assign<<<ob, 256, 0, st>>>(d_vals.as<Fr>(), d_table.as<uint32_t>(), d_slot.as<uint32_t>(),
                                          d_dpos.as<uint32_t>(), d_where.as<uint32_t>(), n_other, a.cidx[0], a.cidx[1],
                                          a.cidx[2], d_coef.as<Fr>());
    }
    FB_CUDA(cudaStreamSynchronize(st));
    FB_CUDA(cudaGetLastError());
    t_kern = (float)((now_s() - t0 - t_alloc) * 1e3);
    t0 = now_s();
    for (int m = 0; m < 3; m++) {
      out.col[m].resize(nnz[m]);
      out.cidx[m].resize(nnz[m]);
      FB_CUDA(cudaMemcpyAsync(out.col[m].data(), d_col[m].p, nnz[m] * 4, cudaMemcpyDeviceToHost, st));
      FB_CUDA(cudaMemcpyAsync(out.cidx[m].data(), d_cidx[m].p, nnz[m] * 4, cudaMemcpyDeviceToHost, st));
    }
    out.coef.resize(n_coef);
    if (n_coef) FB_CUDA(cudaMemcpyAsync(out.coef.data(), d_coef.p, (size_t)n_coef * sizeof(Fr), cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaStreamSynchronize(st));
    t_down = (float)((now_s() - t0) * 1e3);
    out.n_gates = n_gates;
    if (times_ms) {
      times_ms[0] = (float)(t_frame * 1e3); times_ms[1] = t_up; times_ms[2] = t_kern; times_ms[3] = t_down;
      times_ms[5] = (float)(t_alloc * 1e3);
    }
    return FB_OK;
  }
  // no gate parses
  for (int m = 0; m < 3; m++) { out.rowptr[m].assign(1, 0); out.col[m].clear(); out.cidx[m].clear(); }
  out.coef.clear();
  out.n_gates = 0;
  if (times_ms) { times_ms[0] = (float)(t_frame * 1e3); times_ms[1] = t_up; times_ms[2] = 0; times_ms[3] = 0; times_ms[5] = 0; }
  return FB_OK;
}

}  // namespace fb
