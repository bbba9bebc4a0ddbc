// Host-side handle for the Fr evaluation domain of size 2^k (see ntt.cu).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>

#include "ff.cuh"

namespace fb {

constexpr int NTT_LOG_TILE = 12;   // 4096 elements = 128 KiB of shared memory per CTA
constexpr int NTT_MIN_LO = 3;      // strided passes move >= 8 consecutive elements (256 B)
constexpr int NTT_THREADS = 512;

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
  void ifft_then_coset_fft(Fr* x, cudaStream_t st) const;
  void pointwise_then_icoset_fft(Fr* a, const Fr* b, const Fr* c, cudaStream_t st) const;
  void transform(Fr* x, Fr* scratch, int kind, cudaStream_t st) const;
  void bitrev(Fr* dst, const Fr* src, cudaStream_t st) const;
};

}  // namespace fb
