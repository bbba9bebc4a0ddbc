// Pippenger bucket multi-scalar multiplication for BN254 G1 / G2 on sm_100a.
//
// Replaces bellman_ce's `multiexp` (un-vendored; eight call sites inside
// create_random_proof, reached from fawkes-crypto/src/backend/bellman_groth16/
// prover.rs:80).  The result sum_i s_i * P_i is a unique group
// element, so a different window size / digit encoding / coordinate system than
// bellman's (c = ln n unsigned windows, Jacobian) gives byte-identical affine output.
//
// Pipeline (all on one stream):
//   0. (key load)      window tables: 2^(c w) P_i for every window w next to the bases (k_build_window_table),
//                      so that all windows share ONE set of 2^(c-1) buckets and c can grow by ~4 bits
//   1. k_digits<0>     canonical scalar (one Montgomery multiply by 1) -> signed base-2^c digits d in
//                      [-2^(c-1), 2^(c-1)], histogram of bucket = |d| - 1 (+ window * 2^(c-1) without tables)
//   2. k_scan_*        exclusive prefix sum of the histogram (bucket offsets)
//   3. k_digits<1>     counting-sort placement of (w * n + i | sign << 31)
//   4. k_accumulate    equal-length tasks of 2^task_log sorted entries per thread, XYZZ mixed adds, a partial
//                      flushed at every bucket change; k_bucket_gather / k_bucket_heavy fold the partials
//   5. k_bucket_segments / k_segment_bits / k_fold_parts   per-segment short running sums, then the segment
//                      sums combined bit-wise into V[p] (weight 2^p)
//   6. (host)          sum_p 2^p V[p] by Horner on a CPU core: msm_horner_host
// Bases are resident in HBM in Montgomery affine form, 64 B (G1) / 128 B (G2).
#pragma once
#include <cuda_runtime.h>

#include "ec.cuh"
#include "host_fq.h"

namespace fb {

constexpr int MSM_MIN_TASK_LOG = 4;
constexpr uint32_t MSM_MIN_TASK = 1u << MSM_MIN_TASK_LOG;
constexpr uint32_t MSM_HEAVY = 32;        // partials per bucket handled by one thread
constexpr int MSM_HEAVY_THREADS = 128;
constexpr uint32_t MSM_COOP_SEGMENTS_MAX = 1024;   // up to here the segment sums use eight lanes per segment (latency path; measured: no gain at 4096-8192 segments, where the launch is work-bound)
constexpr uint32_t MSM_WARP_GATHER_MAX_BUCKETS = 4096;  // up to here buckets are folded by a warp each (latency path; measured: at 32 k buckets it already costs a batched prove +35 %)
constexpr int MSM_VBITS = 288;            // >= W*c for every plan (255 + c - 1 <= 274 for c <= 20)

struct MsmPlan {
  uint32_t n = 0;       // number of (scalar, base) pairs
  int c = 0;            // window bits
  int W = 0;            // number of windows (digits per scalar), W*c >= 255
  uint32_t B = 0;       // buckets per window = 2^(c-1)
  int seg_log = 0;      // buckets per reduce segment = 2^seg_log
  int task_log = 4;     // entries per accumulation task = 2^task_log
  // Window-table mode: the base array holds 2^(c w) P_i at index w*n + i for every window w (built once
  // per key, k_build_window_table), so digit w of scalar i is just one more point for the ONE shared
  // bucket set.  The reduction then costs 2^(c-1) buckets instead of W * 2^(c-1), which lets c grow by
  // ~4 bits: 15 -> 12 digits per scalar at n = 2^24.  Costs W x the base memory (HBM is 180 GB).
  bool table = false;
  // Batched MSM (fb_prove_batch on a small key): `sets` independent scalar vectors over the SAME bases, set p
  // reading scalars[p * sstride + ...]; every set gets its own buckets (bucket id = set * B + digit - 1), so one
  // sort, one accumulation and one reduction carry the digits of all the proofs of a batch.
  uint32_t sets = 1;
  uint64_t sstride = 0;
  bool batched_out = false;  // result array holds vbits() entries (set p at p * vbits_per_set()), not MSM_VBITS
  static MsmPlan make(uint32_t n, bool table = false);
  MsmPlan batched(uint32_t nsets, uint64_t stride) const;      // same window, task size for the whole batch
  int task_log_for_bucket_size() const;
  uint64_t entries() const { return (uint64_t)n * W * sets; }  // digits to sort and accumulate (upper bound)
  int wred() const { return (int)sets * (table ? 1 : W); }     // bucket sets to reduce
  int vbits_per_set() const { return (table ? 1 : W) * c; }    // V entries of one scalar vector
  uint32_t nbuckets() const { return (uint32_t)wred() * B; }
  uint32_t nsegs() const { return nbuckets() >> seg_log; }
  int vbits() const { return wred() * c; }                     // V entries the host Horner consumes
  uint64_t table_points() const { return table ? (uint64_t)n * W : n; }
};

// Scratch shared by consecutive MSMs on one stream (sized for the largest plan).
struct MsmScratch {
  uint32_t* hist = nullptr;     // [nbuckets + 1]
  uint32_t* offsets = nullptr;  // [nbuckets + 1]
  uint32_t* cursor = nullptr;   // [nbuckets]
  uint32_t* blocksums = nullptr;
  uint32_t* sorted = nullptr;   // [n * W]
  void* buckets = nullptr;      // [nbuckets] XYZZ (sized for G2)
  void* segR = nullptr;         // [nsegs] XYZZ: per-segment weighted sums
  void* segS = nullptr;         // [nsegs] XYZZ: per-segment plain sums
  void* winsum = nullptr;       // [(1 + sbits) * wred * parts] XYZZ: per-CTA partial bit sums (k_segment_bits)
  uint32_t* ntasks = nullptr;   // [nbuckets + 1]
  uint32_t* task_off = nullptr; // [nbuckets + 1]
  void* partials = nullptr;     // [cap_tasks] XYZZ
  uint32_t* heavy = nullptr;    // [0] = count, then bucket ids
  size_t cap_entries = 0, cap_buckets = 0, cap_tasks = 0;
  int alloc(const MsmPlan* plans, int count, bool need_g2);
  void release();
};

// scalars: Montgomery Fr, addressed as scalars[map ? map[i] : i]; bases: affine Montgomery.
// result: out[MSM_VBITS] on device (out[plan.vbits()] for a batched plan, set p at out + p * vbits_per_set()),
// the per-bit sums V[p] with  MSM = sum_p 2^p V[p]  (finish with
// msm_horner_host after copying them to the host).  Steps 1-3 are skipped when reuse_sort is set
// (same scalars as the previous call on this scratch: B_g2 then B_g1).
int msm_g1(const G1Affine* bases, const Fr* scalars, const uint32_t* map, const MsmPlan& plan,
           MsmScratch& s, G1XYZZ* out, bool reuse_sort, cudaStream_t st);
int msm_g2(const G2Affine* bases, const Fr* scalars, const uint32_t* map, const MsmPlan& plan,
           MsmScratch& s, G2XYZZ* out, bool reuse_sort, cudaStream_t st);
// Fill tab[w*n + i] = 2^(c w) tab[i] for w = 1..W-1 (tab[0..n) holds the bases); plan.table must be set.
int msm_build_table_g1(G1Affine* tab, const MsmPlan& plan, cudaStream_t st);
int msm_build_table_g2(G2Affine* tab, const MsmPlan& plan, cudaStream_t st);

// Host side of step 6: Horner over the bit sums, on 64-bit-limb host arithmetic.
template <class HF, class DF>
inline XYZZ<HF> msm_horner_host(const XYZZ<DF>* V, int nbits) {
  XYZZ<HF> acc = XYZZ<HF>::inf();
  for (int p = nbits - 1; p >= 0; p--) {
    if (!acc.is_inf()) acc = dbl(acc);
    if (!V[p].is_inf()) {
      XYZZ<HF> t{HF::from(V[p].x), HF::from(V[p].y), HF::from(V[p].zz), HF::from(V[p].zzz)};
      acc = add(acc, t);
    }
  }
  return acc;
}

}  // namespace fb
