// Host-side handle for the Fr evaluation domain of size 2^k (see ntt.cu).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>

#include "ff.cuh"

namespace fb {

// 1024 elements = 32 KiB of shared memory per CTA, 256 threads: five CTAs per SM overlap each other's load,
// compute and store phases.  Measured H pipeline at 2^24: 35.7 ms with 128 KiB tiles x 512 threads (one CTA
// per SM), 30.0 with 64 KiB x 512, 28.4 with 64 KiB x 256, 27.3 with 32 KiB x 256, 28.6 with 16 KiB x 256.
constexpr int NTT_LOG_TILE = 10;
constexpr int NTT_MIN_LO = 3;      // strided passes move >= 8 consecutive elements (256 B)
constexpr int NTT_THREADS = 256;

// Transport of the distributed transform: every rank sends chunk r (count elements) of each of
// `narrays` send buffers to rank r and receives chunk r of each recv buffer from rank r.
struct NttExchange {
  virtual ~NttExchange() {}
  virtual int all_to_all(const Fr* const* send, Fr* const* recv, int narrays, uint64_t count,
                         cudaStream_t st) = 0;
  // A second stream (+ events) on which exchanges may run beside the transforms of the other arrays; nullptr =
  // the transport wants everything on the compute stream (the single-process test stand-in).
  virtual cudaStream_t side_stream() { return nullptr; }
  virtual cudaEvent_t event(int i) { (void)i; return nullptr; }   // i in [0, 12)
};

struct NttDomain {
  int k = 0;
  Fr omega, minv, k1, k2;
  Fr* tab_plain = nullptr;    // w^(j 2^t)
  Fr* tab_coset = nullptr;    // (g w^j)^(2^t)
  Fr* tab_icoset = nullptr;   // (g w^j)^(-2^t)
  int plan_bm = 0;            // bits handled by the contiguous pass
  int n_strided = 0;
  int pass_lb[8], pass_b[8];

  int init(int k, cudaStream_t st);
  void destroy();
  // batch > 1: the same transform on `batch` arrays, bstride elements apart (fb_prove_batch)
  void ifft_then_coset_fft(Fr* x, cudaStream_t st, unsigned batch = 1, uint64_t bstride = 0) const;
  void pointwise_then_icoset_fft(Fr* a, const Fr* b, const Fr* c, cudaStream_t st, unsigned batch = 1,
                                 uint64_t bstride = 0) const;
  void transform(Fr* x, Fr* scratch, int kind, cudaStream_t st) const;
  void bitrev(Fr* dst, const Fr* src, cudaStream_t st) const;
  // distributed H pipeline over G = 2^g ranks (see ntt.cu); ev/tmp are local arrays of 2^(k-g)
  bool dist_supported(int g) const;
  void dist_plan(int g, int* n_pass, int* lb, int* b) const;
  int dist_h_pipeline(Fr* const ev[3], Fr* const tmp[3], int g, int rank, NttExchange* xch,
                      cudaStream_t st) const;
};

}  // namespace fb
