// See msm.cuh for the design and the reference interface this replaces.
#include "msm.cuh"
#include "internal.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace fb {

// ------------------------------------------------------------------ plan ---
MsmPlan MsmPlan::make(uint32_t n, bool table) {
  MsmPlan p;
  p.n = n;
  p.table = table;
  int lg = 0;
  while ((2u << lg) <= n && lg < 31) lg++;
  int c;
  if (table) {
    // One bucket set for all windows: n*W mixed adds to accumulate, and sort + reduction work per bucket worth ~15
    // mixed adds (2 ns per bucket against 0.135 ns per add; measured sweeps: profiles/r01_window_sweep.txt, and this
    // round profiles/r02_window_sweep_2e21.txt, r02_window_sweep_2e22.txt: 2^22 points c = 17 / 19 / 20 -> 66.5 / 69.0 /
    // 63.5 ms).  Windows whose top digit has only a few bits are skipped: all n top digits then land in a handful of
    // buckets, thousands of atomics per counter in the digit sort (c = 21 / 23 at 2^24 cost +30 / +70 ms; c = 19 at
    // 2^22 doubles the sort).  Small circuits (at most 4096 entries per top bucket, at least 3 top bits) keep them: at
    // n = 8191 the rule left only c = 8 -- 32 digits per scalar, every bucket "heavy" -- where c = 10 needs 26
    // (profiles/r02_small_circuit_window_sweep.txt).
    double best = 1e300;
    c = 0;
    for (int cand = std::max(4, lg - 6); cand <= std::min(22, std::max(4, lg)); cand++) {
      const int Wc = (255 + cand - 1) / cand;
      const int top_bits = 254 - (Wc - 1) * cand;
      if (2 * top_bits < cand && (top_bits < 3 || (n >> top_bits) > 4096u)) continue;
      const double cost = (double)n * Wc + 15.0 * (double)(1u << (cand - 1));
      if (cost < best) { best = cost; c = cand; }
    }
    if (c == 0) c = std::max(4, std::min(22, lg - 3));
  } else {
    // accumulation costs n*W mixed adds, reduction ~2.6 * W * 2^(c-1) full adds
    c = lg - 4;
    if (c > 17) c = 17;  // measured at 2^24: c = 17 (W = 15) beats 18..20, whose bucket arrays fall out of L2
  }
  if (c < 4) c = 4;
  if (const char* e = getenv(table ? "FB_MSM_TABLE_C" : "FB_MSM_C")) {
    int v = atoi(e);
    if (v >= 4 && v <= 24) c = v;
  }
  p.c = c;
  p.W = (255 + c - 1) / c;
  p.B = 1u << (c - 1);
  p.seg_log = 3;   // K = 8 buckets per reduce segment: 16 dependent adds deep (c >= 4)
  // task size: enough tasks to fill the chip, at most 256 adds per thread
  uint64_t want = ((uint64_t)n * p.W) >> 18;
  int tl = 4;
  while (tl < 8 && (1ull << tl) < want) tl++;
  if (tl < MSM_MIN_TASK_LOG) tl = MSM_MIN_TASK_LOG;
  if (const char* e = getenv("FB_MSM_TASK_LOG")) {
    int v = atoi(e);
    if (v >= MSM_MIN_TASK_LOG && v <= 10) tl = v;
  }
  p.task_log = tl;
  return p;
}

MsmPlan MsmPlan::batched(uint32_t nsets, uint64_t stride) const {
  MsmPlan p = *this;
  p.sets = nsets;
  p.sstride = stride;
  p.batched_out = true;
  uint64_t want = p.entries() >> 18;
  int tl = 4;
  while (tl < 8 && (1ull << tl) < want) tl++;
  p.task_log = std::max(tl, p.task_log_for_bucket_size());
  return p;
}

// Tasks long enough that an average bucket is covered by at most ~24 of them (MSM_HEAVY = 32 partial sums is where a
// bucket goes onto the "heavy" path): with few, big buckets short tasks would make every bucket heavy.
int MsmPlan::task_log_for_bucket_size() const {
  const uint64_t per_bucket = (uint64_t)n * W / std::max<uint32_t>(table ? B : B * (uint32_t)W, 1u);
  int tl = MSM_MIN_TASK_LOG;
  while (tl < 8 && (24ull << tl) < per_bucket) tl++;
  return tl;
}

// segments one k_segment_bits CTA folds (per (window, bit) job the segment range is cut into parts)
constexpr int MSM_BITS_PART_LOG = 11;
static inline int bits_parts(const MsmPlan& p) {
  const int sbits = p.c - 1 - p.seg_log;
  return sbits > MSM_BITS_PART_LOG ? 1 << (sbits - MSM_BITS_PART_LOG) : 1;
}

int MsmScratch::alloc(const MsmPlan* plans, int count, bool need_g2) {
  // capacity = max over the plans that will actually run on this scratch
  uint64_t ent = 1, bk = 1, tk = 1, vp = 1;
  for (int i = 0; i < count; i++) {
    const MsmPlan& p = plans[i];
    if (p.n == 0) continue;
    ent = std::max<uint64_t>(ent, p.entries());
    bk = std::max<uint64_t>(bk, p.nbuckets());
    tk = std::max<uint64_t>(tk, (p.entries() >> p.task_log) + p.nbuckets() + 2);
    vp = std::max<uint64_t>(vp, (uint64_t)p.wred() * (p.c - p.seg_log) * bits_parts(p));
  }
  cap_entries = ent;
  cap_buckets = bk;
  size_t psz = need_g2 ? sizeof(G2XYZZ) : sizeof(G1XYZZ);
  if (cudaMalloc(&hist, (bk + 1) * 4) != cudaSuccess) return -1;
  if (cudaMalloc(&offsets, (bk + 1) * 4) != cudaSuccess) return -1;
  if (cudaMalloc(&cursor, (bk + 1) * 4) != cudaSuccess) return -1;
  if (cudaMalloc(&blocksums, ((bk + 1023) / 1024 + 1) * 4) != cudaSuccess) return -1;
  if (cudaMalloc(&sorted, std::max<uint64_t>(ent, 1) * 4) != cudaSuccess) return -1;
  if (cudaMalloc(&buckets, bk * psz) != cudaSuccess) return -1;
  if (cudaMalloc(&segR, (bk / 2 + 1) * psz) != cudaSuccess) return -1;  // per-segment weighted sums (K >= 2)
  if (cudaMalloc(&segS, (bk / 2 + 1) * psz) != cudaSuccess) return -1;  // per-segment plain sums
  if (cudaMalloc(&winsum, vp * psz) != cudaSuccess) return -1;
  // task decomposition of the bucket runs (load balancing under skewed digits)
  cap_tasks = tk + 1;
  if (cudaMalloc(&ntasks, (bk + 1) * 4) != cudaSuccess) return -1;
  if (cudaMalloc(&task_off, (bk + 1) * 4) != cudaSuccess) return -1;
  if (cudaMalloc(&partials, cap_tasks * psz) != cudaSuccess) return -1;
  if (cudaMalloc(&heavy, (cap_tasks / MSM_HEAVY + 2) * 4) != cudaSuccess) return -1;
  return 0;
}

void MsmScratch::release() {
  cudaFree(hist); cudaFree(offsets); cudaFree(cursor); cudaFree(blocksums); cudaFree(sorted);
  cudaFree(buckets); cudaFree(segR); cudaFree(segS); cudaFree(winsum);
  cudaFree(ntasks); cudaFree(task_off); cudaFree(partials); cudaFree(heavy);
  ntasks = task_off = heavy = nullptr; partials = nullptr;
  hist = offsets = cursor = blocksums = sorted = nullptr;
  buckets = segR = segS = winsum = nullptr;
}

// ---------------------------------------------------------------- digits ---
__device__ __forceinline__ uint32_t window_bits(const uint32_t* s, int pos, int c) {
  const int limb = pos >> 5, sh = pos & 31;
  if (limb >= 8) return 0;
  uint64_t v = s[limb];
  if (limb + 1 < 8) v |= (uint64_t)s[limb + 1] << 32;
  return (uint32_t)(v >> sh) & ((1u << c) - 1);
}

// SCATTER=false: histogram.  SCATTER=true: place entries using cursor (pre-loaded with offsets).
template <bool SCATTER>
__global__ void k_digits(const Fr* __restrict__ scalars, const uint32_t* __restrict__ map,
                         uint32_t n, int c, int W, uint32_t B, bool table, uint32_t sets, uint64_t sstride,
                         uint32_t* __restrict__ counter, uint32_t* __restrict__ sorted) {
  const uint64_t total = (uint64_t)n * sets;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t set = sets == 1 ? 0u : (uint32_t)(j / n);
    const uint32_t i = (uint32_t)(j - (uint64_t)set * n);
    Fr s = from_mont(scalars[set * sstride + (map ? map[i] : i)]);
    uint32_t carry = 0;
    for (int w = 0; w < W; w++) {
      uint32_t v = window_bits(s.v, w * c, c) + carry;
      uint32_t neg = 0, mag = v;
      if (v > B) {  // v in (2^(c-1), 2^c] -> negative digit v - 2^c, carry 1
        mag = (1u << c) - v;
        neg = 1;
        carry = 1;
      } else {
        carry = 0;
      }
      if (mag != 0) {
        // table mode: every window shares one bucket set and digit w selects the point 2^(c w) P_i
        uint32_t bucket = (table ? set : set * (uint32_t)W + (uint32_t)w) * B + mag - 1;
        uint32_t pos = atomicAdd(&counter[bucket], 1u);
        if (SCATTER) sorted[pos] = (table ? (uint32_t)w * n + i : i) | (neg << 31);
      }
    }
  }
}

// ------------------------------------------------------------------ scan ---
__global__ void k_scan_block(const uint32_t* in, uint32_t* out, uint32_t* __restrict__ blocksums, uint32_t n) {
  __shared__ uint32_t wsum[32];
  const uint32_t i = blockIdx.x * 1024 + threadIdx.x;
  uint32_t v = i < n ? in[i] : 0;
  uint32_t x = v;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) wsum[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = wsum[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += y;
    }
    wsum[lane] = w;
  }
  __syncthreads();
  uint32_t incl = x + (wid ? wsum[wid - 1] : 0);
  if (i < n) out[i] = incl - v;  // exclusive within block
  if (threadIdx.x == 1023) blocksums[blockIdx.x] = incl;
}
__global__ void k_scan_sums(uint32_t* blocksums, uint32_t nb) {  // single thread block, serial over chunks
  __shared__ uint32_t carry_s;
  __shared__ uint32_t wsum[32];
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nb; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    uint32_t v = i < nb ? blocksums[i] : 0, x = v;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) wsum[wid] = x;
    __syncthreads();
    if (wid == 0) {
      uint32_t w = wsum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += y;
      }
      wsum[lane] = w;
    }
    __syncthreads();
    uint32_t incl = x + (wid ? wsum[wid - 1] : 0) + carry_s;
    if (i < nb) blocksums[i] = incl - v;  // exclusive
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = incl;
    __syncthreads();
  }
}
__global__ void k_scan_add(uint32_t* __restrict__ out, uint32_t* __restrict__ cursor,
                           const uint32_t* __restrict__ blocksums, uint32_t n) {
  const uint32_t i = blockIdx.x * 1024 + threadIdx.x;
  if (i < n) {
    uint32_t v = out[i] + blocksums[blockIdx.x];
    out[i] = v;
    if (cursor) cursor[i] = v;
  }
}

// ------------------------------------------------------------ accumulate ---
#ifndef FB_G2_ACC_CTAS
#define FB_G2_ACC_CTAS 2
#endif
#define MSM_ACC_MIN_CTAS(F) (sizeof(F) == sizeof(Fq) ? 4 : FB_G2_ACC_CTAS)
// Equal-length tasks: thread t accumulates the sorted entries [t*T, (t+1)*T), T = 2^task_log, and
// flushes a partial sum whenever the bucket changes.  Every thread performs exactly T mixed adds
// (no divergence in trip count, no matter how skewed the digits are: a bucket holding a third of
// the points is simply spread over thousands of threads).  The partial of (bucket b, thread t)
// lives at index b + t: bucket b owns the contiguous slots b + floor(start_b / T) .. b +
// floor((end_b - 1) / T), so no second scan is needed to find them.
#ifndef MSM_LD64_DEFAULT
#define MSM_LD64_DEFAULT 1
#endif
static int msm_ld64() {  // FB_MSM_LD64=0 restores plain loads (A/B measurements)
  static const int v = [] { const char* e = getenv("FB_MSM_LD64"); return e ? atoi(e) : MSM_LD64_DEFAULT; }();
  return v;
}

// A base is gathered from a random place of a multi-gigabyte array.  A plain load lets L2 fetch a whole
// 128-byte line for a 64-byte G1 point (ncu at 2^24: 29.3 GB of DRAM reads for 14.8 GB of points);
// the .L2::64B prefetch-size hint asks for the two sectors that are used.
template <class F>
__device__ __forceinline__ Affine<F> load_base(const Affine<F>* __restrict__ p, int ld64) {
  if (sizeof(Affine<F>) == 64 && ld64) {
    Affine<F> r;
    uint4* d = reinterpret_cast<uint4*>(&r);
    const uint4* s = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < 4; i++)
      asm volatile("ld.global.nc.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(d[i].x), "=r"(d[i].y), "=r"(d[i].z), "=r"(d[i].w)
                   : "l"(s + i));
    return r;
  }
  return *p;
}

template <class F>
__global__ void __launch_bounds__(128, MSM_ACC_MIN_CTAS(F))
k_accumulate(const Affine<F>* __restrict__ bases, const uint32_t* __restrict__ sorted,
             const uint32_t* __restrict__ offsets, uint32_t nb, int task_log,
             XYZZ<F>* __restrict__ partials, int ld64) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t total = offsets[nb];
  const uint32_t start = t << task_log;
  if (start >= total) return;
  const uint32_t end = min(start + (1u << task_log), total);
  // bucket containing entry `start`: largest b with offsets[b] <= start
  uint32_t lo = 0, hi = nb;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (offsets[mid] <= start) lo = mid; else hi = mid;
  }
  uint32_t b = lo;
  uint32_t bend = offsets[b + 1];
  XYZZ<F> acc = XYZZ<F>::inf();
  // G1: the next base is fetched into registers while the current add runs.  G2: a second 128-byte point
  // in registers spills (255 registers either way), so the next point is only prefetched into L2 and
  // loaded at the top of its own iteration -- an L2 hit against a 17k-cycle add.
  constexpr bool REG_PREFETCH = sizeof(F) == sizeof(Fq);
  uint32_t e = sorted[start];
  Affine<F> nxt;
  if (REG_PREFETCH) nxt = load_base(bases + (e & 0x7fffffffu), ld64);
  for (uint32_t p = start; p < end; p++) {
    if (p >= bend) {  // bucket boundary: flush and move on (empty buckets are skipped)
      partials[b + t] = acc;
      acc = XYZZ<F>::inf();
      do { b++; bend = offsets[b + 1]; } while (p >= bend);
    }
    Affine<F> cur;
    const uint32_t sign = e >> 31;
    if (REG_PREFETCH) {
      cur = nxt;
      if (p + 1 < end) {
        e = sorted[p + 1];
        nxt = load_base(bases + (e & 0x7fffffffu), ld64);
      }
    } else {
      const Affine<F>* src = bases + (e & 0x7fffffffu);
      if (p + 1 < end) {
        e = sorted[p + 1];
        const Affine<F>* nsrc = bases + (e & 0x7fffffffu);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nsrc));
      }
      cur = *src;
    }
    if (sign) cur.y = neg(cur.y);
    acc = add_mixed(acc, cur);
  }
  partials[b + t] = acc;
}

// bucket = sum of its partials; buckets with more than MSM_HEAVY partials are queued
template <class F>
__global__ void __launch_bounds__(128)
k_bucket_gather(const XYZZ<F>* __restrict__ partials, const uint32_t* __restrict__ offsets,
                uint32_t nb, int task_log, XYZZ<F>* __restrict__ buckets, uint32_t* __restrict__ heavy) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const uint32_t s = offsets[b], e = offsets[b + 1];
  if (s == e) { buckets[b] = XYZZ<F>::inf(); return; }
  const uint32_t t0 = s >> task_log, t1 = (e - 1) >> task_log;
  if (t1 - t0 + 1 > MSM_HEAVY) {
    heavy[1 + atomicAdd(&heavy[0], 1u)] = b;
    return;
  }
  XYZZ<F> acc = partials[b + t0];
  for (uint32_t t = t0 + 1; t <= t1; t++) acc = add_cold(acc, partials[b + t]);
  buckets[b] = acc;
}

// Latency variant for small problems (a single small prove is a chain of dependent adds, not a throughput problem):
// one WARP per bucket, the partials strided over the lanes and folded with five shuffle steps, so a bucket with 26
// partial sums is 6 dependent adds deep instead of 26.  Buckets with more than `heavy_min` partials are queued.
template <class F>
__device__ __forceinline__ XYZZ<F> shfl_down_xyzz(const XYZZ<F>& v, int d) {
  XYZZ<F> r;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* o = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(XYZZ<F>) / 4); i++) o[i] = __shfl_down_sync(0xffffffffu, s[i], d);
  return r;
}
template <class F>
__global__ void __launch_bounds__(128)
k_bucket_gather_warp(const XYZZ<F>* __restrict__ partials, const uint32_t* __restrict__ offsets, uint32_t nb, int task_log,
                     uint32_t heavy_min, XYZZ<F>* __restrict__ buckets, uint32_t* __restrict__ heavy) {
  const uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= nb) return;  // warp-uniform
  const uint32_t s = offsets[b], e = offsets[b + 1];
  if (s == e) { if (lane == 0) buckets[b] = XYZZ<F>::inf(); return; }
  const uint32_t t0 = s >> task_log, t1 = (e - 1) >> task_log;
  if (t1 - t0 + 1 > heavy_min) {
    if (lane == 0) heavy[1 + atomicAdd(&heavy[0], 1u)] = b;
    return;
  }
  XYZZ<F> acc = XYZZ<F>::inf();
  for (uint32_t t = t0 + lane; t <= t1; t += 32) acc = add_cold(acc, partials[b + t]);
  if (t1 > t0) {  // warp-uniform: a single partial needs no fold
#pragma unroll 1
    for (int d = 16; d > 0; d >>= 1) {
      const XYZZ<F> o = shfl_down_xyzz(acc, d);
      acc = add_cold(acc, o);
    }
  }
  if (lane == 0) buckets[b] = acc;
}

// one CTA per queued bucket: strided partial sums then a shared-memory tree
template <class F>
__global__ void __launch_bounds__(MSM_HEAVY_THREADS)
k_bucket_heavy(const XYZZ<F>* __restrict__ partials, const uint32_t* __restrict__ offsets,
               int task_log, const uint32_t* __restrict__ heavy, XYZZ<F>* __restrict__ buckets) {
  __shared__ XYZZ<F> sh[MSM_HEAVY_THREADS];
  const uint32_t count = heavy[0];
  for (uint32_t i = blockIdx.x; i < count; i += gridDim.x) {
    const uint32_t b = heavy[1 + i];
    const uint32_t t0 = offsets[b] >> task_log, t1 = (offsets[b + 1] - 1) >> task_log;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t t = t0 + threadIdx.x; t <= t1; t += MSM_HEAVY_THREADS) acc = add_cold(acc, partials[b + t]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = MSM_HEAVY_THREADS / 2; s > 0; s >>= 1) {
      if ((int)threadIdx.x < s) sh[threadIdx.x] = add_cold(sh[threadIdx.x], sh[threadIdx.x + s]);
      __syncthreads();
    }
    if (threadIdx.x == 0) buckets[b] = sh[0];
    __syncthreads();
  }
}

// ---------------------------------------------------------------- reduce ---
// sum_w 2^(c w) sum_j (j+1) B[w][j]  in three parallel steps (no long serial running sum):
//  1. k_bucket_segments: segments of K = 2^seg_log buckets, one thread each, short running sums:
//        R_s = sum_t (t+1) B[sK+t],  S_s = sum_t B[sK+t]      so that  sum_j (j+1) B_j = sum_s R_s + K sum_s s S_s
//  2. k_segment_bits: sum_s s S_s is evaluated bit-wise, V[c w + seg_log + k] = sum of S_s with bit k of
//     s set (one CTA per (w, k), tree), and V[c w] = sum_s R_s
//  3. the caller copies V (<= MSM_VBITS points) to the host and evaluates sum_p 2^p V[p] by Horner
//     there (msm_horner_host): the ~255 dependent doublings cost 0.2 ms on a CPU core against
//     1.3 ms (G1) / 4 ms (G2) for one GPU thread, and they overlap the remaining device work.
// Cost ~2 adds per bucket, so large windows (few digits per scalar) stay cheap to reduce.
template <class F>
__global__ void __launch_bounds__(128)
k_bucket_segments(const XYZZ<F>* __restrict__ buckets, uint32_t nsegs, int seg_log,
                  XYZZ<F>* __restrict__ segR, XYZZ<F>* __restrict__ segS) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsegs) return;
  const XYZZ<F>* bk = buckets + ((uint64_t)s << seg_log);
  XYZZ<F> run = XYZZ<F>::inf(), tot = XYZZ<F>::inf();
  for (int t = (1 << seg_log) - 1; t >= 0; t--) {
    run = add(run, bk[t]);   // inlined: 2K dependent adds are this kernel's critical path
    tot = add_cold(tot, run);
  }
  segR[s] = tot;
  segS[s] = run;
}

// Latency variant for at most MSM_COOP_SEGMENTS_MAX segments (a single small prove: configs[0] 1.6 -> 1.3 ms):
// EIGHT lanes per segment of 8 buckets.  Suffix sums run_t = sum_{j >= t} B_j by three shuffle steps,
// then R = sum_t run_t (every B_j is counted j + 1 times) by three more, S = run_0: 6 dependent adds instead of 16,
// for 3x the lane-adds -- these launches are far too small to be throughput-bound.
template <class F>
__global__ void __launch_bounds__(128)
k_bucket_segments_coop(const XYZZ<F>* __restrict__ buckets, uint32_t nsegs, XYZZ<F>* __restrict__ segR,
                       XYZZ<F>* __restrict__ segS) {
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t s = gid >> 3, t = gid & 7;
  XYZZ<F> run = s < nsegs ? buckets[((uint64_t)s << 3) + t] : XYZZ<F>::inf();
#pragma unroll 1
  for (int d = 1; d < 8; d <<= 1) {
    XYZZ<F> o;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&run);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(XYZZ<F>) / 4); i++) dst[i] = __shfl_down_sync(0xffffffffu, src[i], d, 8);
    if (t + d < 8) run = add_cold(run, o);
  }
  XYZZ<F> tot = run;
#pragma unroll 1
  for (int d = 4; d > 0; d >>= 1) {
    XYZZ<F> o;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&tot);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(XYZZ<F>) / 4); i++) dst[i] = __shfl_down_sync(0xffffffffu, src[i], d, 8);
    if (t < (uint32_t)d) tot = add_cold(tot, o);
  }
  if (t == 0 && s < nsegs) {
    segR[s] = tot;
    segS[s] = run;
  }
}

// grid = wred * (1 + sbits) * parts CTAs, sbits = bits of the segment index.  Job (w, 0): V[c w] = sum_s R_s;
// job (w, 1 + k): V[c w + seg_log + k] = sum of S_s over segments whose index has bit k set.  A job's
// segment range is cut into `parts` CTAs (window-table plans have one window of up to 2^18 segments);
// with parts > 1 the CTA sums go to Vpart and k_fold_parts adds them up.
template <class F, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_segment_bits(const XYZZ<F>* __restrict__ segR, const XYZZ<F>* __restrict__ segS, int c, int seg_log,
               int sbits, int parts, XYZZ<F>* __restrict__ V, XYZZ<F>* __restrict__ Vpart) {
  __shared__ XYZZ<F> sh[THREADS];
  const int per_w = 1 + sbits;
  const int job = blockIdx.x / parts, part = blockIdx.x % parts;
  const int w = job / per_w, which = job % per_w;
  const uint32_t segs = 1u << sbits;
  XYZZ<F> acc = XYZZ<F>::inf();
  if (which == 0) {
    const XYZZ<F>* R = segR + (uint64_t)w * segs;
    const uint32_t len = segs / parts, lo = part * len;
    for (uint32_t s = lo + threadIdx.x; s < lo + len; s += THREADS) acc = add_cold(acc, R[s]);
  } else {
    const int k = which - 1;
    const XYZZ<F>* S = segS + (uint64_t)w * segs;
    const uint32_t low_mask = (1u << k) - 1;
    const uint32_t len = (segs >> 1) / parts, lo = part * len;
    for (uint32_t t = lo + threadIdx.x; t < lo + len; t += THREADS) {
      const uint32_t sidx = ((t & ~low_mask) << 1) | (1u << k) | (t & low_mask);  // insert a 1 at bit k
      acc = add_cold(acc, S[sidx]);
    }
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int st = THREADS / 2; st > 0; st >>= 1) {
    if ((int)threadIdx.x < st) sh[threadIdx.x] = add_cold(sh[threadIdx.x], sh[threadIdx.x + st]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (parts == 1) V[c * w + (which == 0 ? 0 : seg_log + which - 1)] = sh[0];
    else Vpart[(uint64_t)job * parts + part] = sh[0];
  }
}

// one warp per job: V[...] = sum of the job's `parts` partial sums
template <class F>
__global__ void __launch_bounds__(32)
k_fold_parts(const XYZZ<F>* __restrict__ Vpart, int c, int seg_log, int sbits, int parts, XYZZ<F>* __restrict__ V) {
  __shared__ XYZZ<F> sh[32];
  const int per_w = 1 + sbits;
  const int job = blockIdx.x, w = job / per_w, which = job % per_w;
  XYZZ<F> acc = XYZZ<F>::inf();
  for (int j = threadIdx.x; j < parts; j += 32) acc = add_cold(acc, Vpart[(uint64_t)job * parts + j]);
  sh[threadIdx.x] = acc;
  __syncwarp();
  for (int st = 16; st > 0; st >>= 1) {
    if ((int)threadIdx.x < st) sh[threadIdx.x] = add_cold(sh[threadIdx.x], sh[threadIdx.x + st]);
    __syncwarp();
  }
  if (threadIdx.x == 0) V[c * w + (which == 0 ? 0 : seg_log + which - 1)] = sh[0];
}

// ------------------------------------------------------------ window table ---
// tab[w*n + i] = 2^(c w) * tab[i].  One thread per base: c doublings per window in XYZZ while tracking
// Z (zz = Z^2, zzz = Z^3: each doubling multiplies Z by 2Y), (X, Y) parked in the output slot, then one
// shared inversion for the thread's W-1 points (Montgomery's trick) and x = X/Z^2, y = Y/Z^3 in place.
constexpr int MSM_MAX_W = 64;
template <class F>
__global__ void __launch_bounds__(128)
k_build_window_table(Affine<F>* __restrict__ tab, uint32_t n, int c, int W) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Affine<F> p = tab[i];
  F zs[MSM_MAX_W];  // Z of window w (one() for points at infinity, which are stored as such)
  F X = p.x, Y = p.y, Z = F::one();
  bool inf = p.is_inf();
  for (int w = 1; w < W; w++) {
    for (int j = 0; j < c && !inf; j++) {
      if (Y.is_zero()) { inf = true; break; }
      // dbl-2008-s-1 with zz = Z^2, zzz = Z^3 implicit
      F U = dbl(Y);
      F V = sqr(U);
      F Wc = mul(U, V);
      F S = mul(X, V);
      F X2 = sqr(X);
      F M = add(dbl(X2), X2);
      F X3 = sub(sqr(M), dbl(S));
      Y = sub(mul(M, sub(S, X3)), mul(Wc, Y));
      X = X3;
      Z = mul(Z, U);
    }
    if (inf) {
      tab[(uint64_t)w * n + i] = Affine<F>::inf();
      zs[w] = F::one();
    } else {
      tab[(uint64_t)w * n + i] = Affine<F>{X, Y};
      zs[w] = Z;
    }
  }
  if (W < 2) return;
  // prefix products pr[w] = zs[1] * ... * zs[w], kept in zs by a second array
  F pr[MSM_MAX_W];
  pr[1] = zs[1];
  for (int w = 2; w < W; w++) pr[w] = mul(pr[w - 1], zs[w]);
  F iv = inv_cold(pr[W - 1]);
  for (int w = W - 1; w >= 1; w--) {
    const F iz = w > 1 ? mul(iv, pr[w - 1]) : iv;  // 1 / zs[w]
    iv = mul(iv, zs[w]);
    Affine<F> q = tab[(uint64_t)w * n + i];
    if (q.is_inf()) continue;
    const F iz2 = sqr(iz);
    q.x = mul(q.x, iz2);
    q.y = mul(q.y, mul(iz2, iz));
    tab[(uint64_t)w * n + i] = q;
  }
}

// ----------------------------------------------------------------- driver ---
template <class F>
static int msm_run(const Affine<F>* bases, const Fr* scalars, const uint32_t* map,
                   const MsmPlan& p, MsmScratch& s, XYZZ<F>* out, bool reuse_sort,
                   cudaStream_t st) {
  const uint32_t nb = p.nbuckets();
  if (p.W > MSM_MAX_W || bits_parts(p) > 1024) return -1;
  const size_t vcount = p.batched_out ? (size_t)p.vbits() : (size_t)MSM_VBITS;
  if (p.n == 0) {
    cudaMemsetAsync(out, 0, sizeof(XYZZ<F>) * vcount, st);
    return 0;
  }
  if (p.entries() > s.cap_entries || nb > s.cap_buckets) return -2;
  if (!reuse_sort) {
    kstat_begin(KSTAT_SORT, st);
    // Counting sort on global atomics.  A two-pass LSD radix sort with warp-private shared-memory cursors was built
    // and measured this round (profiles/r02_radix_sort_experiment.txt): its count passes ran in 0.3-0.75 ms at 2^24, but its
    // scatter passes took 7 and 13 ms -- 2368 warps x 1024 bins of open 4/8-byte write streams (78 MB of partial
    // sectors next to a 1.75 GB stream) thrash L2 into read-modify-writes, 90 ms per prove against 18.6 ms here.
    // Beating this path needs tiles sorted in shared memory and written out as full sectors.
    cudaMemsetAsync(s.hist, 0, (size_t)(nb + 1) * 4, st);
    const unsigned dg = (unsigned)std::min<uint64_t>(((uint64_t)p.n * p.sets + 255) / 256, 148 * 16);
    k_digits<false><<<dg, 256, 0, st>>>(scalars, map, p.n, p.c, p.W, p.B, p.table, p.sets, p.sstride, s.hist, nullptr);
    const unsigned sb = (nb + 1 + 1023) / 1024;
    k_scan_block<<<sb, 1024, 0, st>>>(s.hist, s.offsets, s.blocksums, nb + 1);
    k_scan_sums<<<1, 1024, 0, st>>>(s.blocksums, sb);
    k_scan_add<<<sb, 1024, 0, st>>>(s.offsets, s.cursor, s.blocksums, nb + 1);
    k_digits<true><<<dg, 256, 0, st>>>(scalars, map, p.n, p.c, p.W, p.B, p.table, p.sets, p.sstride, s.cursor, s.sorted);
    kstat_end(KSTAT_SORT, st);
  }
  XYZZ<F>* buckets = reinterpret_cast<XYZZ<F>*>(s.buckets);
  XYZZ<F>* segR = reinterpret_cast<XYZZ<F>*>(s.segR);
  XYZZ<F>* segS = reinterpret_cast<XYZZ<F>*>(s.segS);
  XYZZ<F>* partials = reinterpret_cast<XYZZ<F>*>(s.partials);
  const uint64_t N = p.entries();
  const int kind = sizeof(F) == sizeof(Fq) ? KSTAT_ACC_G1 : KSTAT_ACC_G2;
  const uint32_t* offsets = s.offsets;
  const int task_log = p.task_log;
  kstat_begin(kind, st);
  const uint64_t max_threads = (N >> task_log) + 1;
  if (max_threads + nb > s.cap_tasks) return -4;
  cudaMemsetAsync(s.heavy, 0, 4, st);
  k_accumulate<F><<<(unsigned)((max_threads + 127) / 128), 128, 0, st>>>(bases, s.sorted, s.offsets, nb, p.task_log, partials,
                                                                      msm_ld64());
  kstat_end(kind, st);
  count_launch(reuse_sort ? 5 : 10);
  kstat_begin(KSTAT_REDUCE, st);
  if (nb <= MSM_WARP_GATHER_MAX_BUCKETS)
    k_bucket_gather_warp<F><<<(nb * 32 + 127) / 128, 128, 0, st>>>(partials, offsets, nb, task_log, 32u * MSM_HEAVY, buckets, s.heavy);
  else
    k_bucket_gather<F><<<(nb + 127) / 128, 128, 0, st>>>(partials, offsets, nb, task_log, buckets, s.heavy);
  k_bucket_heavy<F><<<148, MSM_HEAVY_THREADS, 0, st>>>(partials, offsets, task_log, s.heavy, buckets);
  const int sbits = p.c - 1 - p.seg_log;  // bits of the segment index within a window
  const uint32_t nsegs = nb >> p.seg_log;
  XYZZ<F>* V = out;  // MSM_VBITS entries (vbits() for a batched plan)
  constexpr int BT = sizeof(F) == sizeof(Fq) ? 256 : 128;
  const int parts = bits_parts(p);
  const int jobs = p.wred() * (1 + sbits);
  cudaMemsetAsync(V, 0, sizeof(XYZZ<F>) * vcount, st);
  if (p.seg_log == 3 && nsegs <= MSM_COOP_SEGMENTS_MAX)
    k_bucket_segments_coop<F><<<(nsegs * 8 + 127) / 128, 128, 0, st>>>(buckets, nsegs, segR, segS);
  else
    k_bucket_segments<F><<<(nsegs + 127) / 128, 128, 0, st>>>(buckets, nsegs, p.seg_log, segR, segS);
  k_segment_bits<F, BT><<<jobs * parts, BT, 0, st>>>(segR, segS, p.c, p.seg_log, sbits, parts, V,
                                                    reinterpret_cast<XYZZ<F>*>(s.winsum));
  if (parts > 1) {
    k_fold_parts<F><<<jobs, 32, 0, st>>>(reinterpret_cast<XYZZ<F>*>(s.winsum), p.c, p.seg_log, sbits, parts, V);
    count_launch(1);
  }
  kstat_end(KSTAT_REDUCE, st);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

int msm_g1(const G1Affine* bases, const Fr* scalars, const uint32_t* map, const MsmPlan& plan,
           MsmScratch& s, G1XYZZ* out, bool reuse_sort, cudaStream_t st) {
  return msm_run<Fq>(bases, scalars, map, plan, s, out, reuse_sort, st);
}
int msm_g2(const G2Affine* bases, const Fr* scalars, const uint32_t* map, const MsmPlan& plan,
           MsmScratch& s, G2XYZZ* out, bool reuse_sort, cudaStream_t st) {
  return msm_run<Fq2>(bases, scalars, map, plan, s, out, reuse_sort, st);
}

template <class F>
static int build_table(Affine<F>* tab, const MsmPlan& p, cudaStream_t st) {
  if (!p.table || p.n == 0 || p.W < 2) return 0;
  if (p.W > MSM_MAX_W) return -1;
  k_build_window_table<F><<<(p.n + 127) / 128, 128, 0, st>>>(tab, p.n, p.c, p.W);
  count_launch(1);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
int msm_build_table_g1(G1Affine* tab, const MsmPlan& plan, cudaStream_t st) { return build_table<Fq>(tab, plan, st); }
int msm_build_table_g2(G2Affine* tab, const MsmPlan& plan, cudaStream_t st) { return build_table<Fq2>(tab, plan, st); }

}  // namespace fb
