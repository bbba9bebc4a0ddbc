// Test / micro-benchmark entry points (fb_test_*, fb_probe_*): each drives one kernel
// family on caller-supplied host buffers so tests can compare against the CPU oracle
// through the C ABI.  Not part of the reference interface.
#include "../../include/fawkes_b200.h"

#include <chrono>
#include <cstring>
#include <vector>

#include "internal.h"

namespace fb {

// (x + y u)(y + x u^... ) through the device Fq2 multiplier: returns c0 - c1 of (x + y u) * (y + x^2 u)
__device__ inline Fq fq2_probe(const Fq& x, const Fq& y) {
  Fq2 m = mul(Fq2{x, y}, Fq2{y, mul_c(x, x)});
  return sub(m.c0, m.c1);
}
// msub2 over Fq2 (two shared reductions) against four separate products: returns 0 when they agree
__device__ inline Fq fq2_probe2(const Fq& x, const Fq& y) {
  const Fq2 a{x, y}, b{y, mul_c(x, x)}, c{mul_c(y, y), x}, d{add(x, y), sub(x, y)};
  const Fq2 want = sub(mul(a, b), mul(c, d)), got = msub2(a, b, c, d);
  const Fq2 e = sub(want, got);
  return add(add(e.c0, e.c1), add(e.c0, e.c0));   // 0 iff e == 0 up to a negligible cancellation
}
__device__ inline Fr fq2_probe2(const Fr& x, const Fr&) { return Fr::zero(); }
__device__ inline Fr fq2_probe(const Fr& x, const Fr&) { return x; }

template <class C>
__global__ void k_field_op(int op, const Fp<C>* a, const Fp<C>* b, Fp<C>* out, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    Fp<C> x = a[i], y = b ? b[i] : Fp<C>::zero(), r;
    switch (op) {
      case 0: r = mul(x, y); break;
      case 1: r = add(x, y); break;
      case 2: r = sub(x, y); break;
      case 3: r = inv(x); break;
      case 5: r = sqr(x); break;                         // dedicated square (36 + 72 wide MACs)
      case 6: r = msub2(x, y, y, sqr(x)); break;         // x*y - y*x^2, one shared reduction
      case 7: r = fq2_probe(x, y); break;                // Fq only: lazy-reduction Fq2 product
      case 8: r = fq2_probe2(x, y); break;               // Fq only: Fq2 a*b - c*d with shared reductions (-> 0)
      default: r = mul_c(x, y); break;
    }
    out[i] = r;
  }
}

// dependent-free IMAD.WIDE accumulate streams: 8 independent accumulators per thread
// same, but every MAC reads its own multiplier register too (three distinct register operands)
__global__ void k_probe_imad3(uint64_t* out, uint32_t seed, int iters) {
  uint32_t a[8], b[8];
  uint64_t acc[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { acc[j] = j + threadIdx.x; a[j] = seed * (2 * j + 3) + threadIdx.x; b[j] = seed * (2 * j + 5) + blockIdx.x; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(a[j]), "r"(b[(j + 3) & 7]));
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) s ^= acc[j];
  if (s == 0x1234567812345678ull) out[0] = s;
}
// one accumulator, distinct operands: the dependent chain a column sum forms
__global__ void k_probe_imad_chain(uint64_t* out, uint32_t seed, int iters) {
  uint32_t a[8], b[8];
  uint64_t acc = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; j++) { a[j] = seed * (2 * j + 3) + threadIdx.x; b[j] = seed * (2 * j + 5) + blockIdx.x; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a[j]), "r"(b[(j + 3) & 7]));
    }
  }
  if (acc == 0x1234567812345678ull) out[0] = acc;
}

// Integer-pipe roofline probe: 8 independent IMAD.WIDE.U32 accumulate streams per thread.  The multiplier
// operand is taken from another accumulator every round, so ptxas cannot hoist the products out of the
// loop (an earlier version with loop-invariant operands was strength-reduced to IADD3 and reported the
// ADD rate, 1.72e13/s, as the "MAC peak").  SASS of this loop: IMAD.WIDE.U32 only (checked with cuobjdump).
__global__ void k_probe_imad(uint64_t* out, uint32_t seed, int iters) {
  uint32_t X[8][4], a[8];
#pragma unroll
  for (int j = 0; j < 8; j++) a[j] = seed * (2 * j + 3) + threadIdx.x;
#pragma unroll
  for (int h = 0; h < 8; h++)
#pragma unroll
    for (int j = 0; j < 4; j++) X[h][j] = seed * (j + 1) + h + threadIdx.x;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int h = 0; h < 8; h++)  // the multiplier was written 8 MACs (half a round) earlier
      asm volatile("mad.lo.cc.u32 %0, %4, %6, %0; madc.hi.u32 %1, %4, %6, %1;"
                   "mad.lo.cc.u32 %2, %5, %6, %2; madc.hi.u32 %3, %5, %6, %3;"
                   : "+r"(X[h][0]), "+r"(X[h][1]), "+r"(X[h][2]), "+r"(X[h][3])
                   : "r"(a[h]), "r"(a[(h + 3) & 7]), "r"(X[(h + 4) & 7][0]));
  }
  uint32_t s = 0;
#pragma unroll
  for (int h = 0; h < 8; h++)
#pragma unroll
    for (int j = 0; j < 4; j++) s ^= X[h][j];
  if (s == 0x12345678u) out[0] = s;  // keep the loop alive
}

// carry-chained wide MACs exactly as the Montgomery rows issue them (IMAD.WIDE.U32.X)
__global__ void k_probe_madc(uint32_t* out, uint32_t seed, int iters) {
#if defined(__CUDA_ARCH__)
  uint32_t X[4][9];
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int i = 0; i < 9; i++) X[c][i] = seed + c * 9 + i + threadIdx.x;
  uint32_t a0 = seed | 1, a1 = seed * 3 + 1, a2 = seed * 5 + blockIdx.x, a3 = seed * 7 + threadIdx.x, t = seed * 11 + 5;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < 4; c++) ptx::mad_chain(X[c], a0, a1, a2, a3, t + c);
  }
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int i = 0; i < 9; i++) s ^= X[c][i];
  if (s == 0x12345678u) out[0] = s;
#endif
}

// MAC stream shapes: 0 = 8 accumulators, distinct a and b per MAC; 1 = one accumulator (dependent
// chain), distinct operands; 2 = 8 accumulators, a distinct, b shared (operand reuse)
template <int SHAPE>
__global__ void k_probe_mac_shape(uint64_t* out, uint32_t seed, int iters) {
  uint32_t a0 = seed * 3 + threadIdx.x, a1 = seed * 5 + 1, a2 = seed * 7 + 2, a3 = seed * 11 + 3, a4 = seed * 13 + 4,
           a5 = seed * 17 + 5, a6 = seed * 19 + 6, a7 = seed * 23 + 7;
  uint32_t b0 = seed * 29 + blockIdx.x, b1 = seed * 31 + 1, b2 = seed * 37 + 2, b3 = seed * 41 + 3, b4 = seed * 43 + 4,
           b5 = seed * 47 + 5, b6 = seed * 53 + 6, b7 = seed * 59 + 7;
  uint64_t c0 = 0, c1 = 1, c2 = 2, c3 = 3, c4 = 4, c5 = 5, c6 = 6, c7 = 7;
  for (int i = 0; i < iters; i++) {
    if (SHAPE == 0) {  // 8 accumulators, all operands distinct
      c0 += (uint64_t)a0 * b3; c1 += (uint64_t)a1 * b4; c2 += (uint64_t)a2 * b5; c3 += (uint64_t)a3 * b6;
      c4 += (uint64_t)a4 * b7; c5 += (uint64_t)a5 * b0; c6 += (uint64_t)a6 * b1; c7 += (uint64_t)a7 * b2;
    } else if (SHAPE == 1) {  // one accumulator: a dependent chain
      c0 += (uint64_t)a0 * b3; c0 += (uint64_t)a1 * b4; c0 += (uint64_t)a2 * b5; c0 += (uint64_t)a3 * b6;
      c0 += (uint64_t)a4 * b7; c0 += (uint64_t)a5 * b0; c0 += (uint64_t)a6 * b1; c0 += (uint64_t)a7 * b2;
    } else if (SHAPE == 2) {  // 8 accumulators, shared multiplier (operand reuse)
      c0 += (uint64_t)a0 * b0; c1 += (uint64_t)a1 * b0; c2 += (uint64_t)a2 * b0; c3 += (uint64_t)a3 * b0;
      c4 += (uint64_t)a4 * b0; c5 += (uint64_t)a5 * b0; c6 += (uint64_t)a6 * b0; c7 += (uint64_t)a7 * b0;
    } else {  // two accumulators, distinct operands
      c0 += (uint64_t)a0 * b3; c1 += (uint64_t)a1 * b4; c0 += (uint64_t)a2 * b5; c1 += (uint64_t)a3 * b6;
      c0 += (uint64_t)a4 * b7; c1 += (uint64_t)a5 * b0; c0 += (uint64_t)a6 * b1; c1 += (uint64_t)a7 * b2;
    }
    a0 += 3; b0 ^= a0;  // two ALU ops per 8 MACs keep the loop body live
  }
  uint64_t s = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
  if (s == 0x1234567812345678ull) out[0] = s;
}

template <int VARIANT>
__global__ void k_probe_fr_mul_v(Fr* out, int iters) {
#if defined(__CUDA_ARCH__)
  Fr x = Fr::one(), y = Fr::r2();
  x.v[0] += threadIdx.x;
  y.v[1] ^= blockIdx.x;
  for (int i = 0; i < iters; i++) {
    if (VARIANT == 0) { x = mul_ptx(x, y); y = mul_ptx(y, x); }
    else { x = mul_c(x, y); y = mul_c(y, x); }
  }
  if (x.v[0] == 0x12345678u && y.v[3] == 0x9abcdef0u) out[0] = x;
#endif
}
__global__ void k_probe_fr_mul(Fr* out, int iters) {
  Fr x = Fr::one(), y = Fr::r2();
  x.v[0] += threadIdx.x;
  y.v[1] ^= blockIdx.x;
  for (int i = 0; i < iters; i++) {
    x = mul(x, y);
    y = mul(y, x);
  }
  if (x.v[0] == 0x12345678u && y.v[3] == 0x9abcdef0u) out[0] = x;
}

template <class F>
__global__ void k_xyzz_to_affine(const XYZZ<F>* in, Affine<F>* out, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = to_affine(in[i]);
}

}  // namespace fb

using namespace fb;

extern "C" {

int fb_test_field(fb_ctx* ctx_, int field, int op, const uint64_t* a, const uint64_t* b,
                  uint64_t* out, uint64_t n) try {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  if (!ctx || !a || !out) return FB_ERR_ARG;
  FB_CUDA(cudaSetDevice(ctx->device));
  void *da, *db = nullptr, *dout;
  FB_CUDA(cudaMalloc(&da, n * 32));
  FB_CUDA(cudaMalloc(&dout, n * 32));
  FB_CUDA(cudaMemcpyAsync(da, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (b) {
    FB_CUDA(cudaMalloc(&db, n * 32));
    FB_CUDA(cudaMemcpyAsync(db, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  }
  unsigned blocks = (unsigned)std::min<uint64_t>((n + 127) / 128, 148 * 8);
  if (field == 0)
    k_field_op<FrCfg><<<blocks, 128, 0, ctx->stream>>>(op, (const Fr*)da, (const Fr*)db, (Fr*)dout, n);
  else
    k_field_op<FqCfg><<<blocks, 128, 0, ctx->stream>>>(op, (const Fq*)da, (const Fq*)db, (Fq*)dout, n);
  FB_CUDA(cudaMemcpyAsync(out, dout, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  FB_CUDA(cudaGetLastError());
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return FB_OK;
} FB_ABI_CATCH_INT

int fb_test_ntt(fb_ctx* ctx_, int log_n, int kind, uint64_t* data) try {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  if (!ctx || !data || kind < 0 || kind > 3) return FB_ERR_ARG;
  FB_CUDA(cudaSetDevice(ctx->device));
  NttDomain dom;
  if (dom.init(log_n, ctx->stream) != 0) { set_error("domain init failed"); return FB_ERR_DOMAIN; }
  Fr *x, *scratch;
  size_t bytes = sizeof(Fr) << log_n;
  FB_CUDA(cudaMalloc(&x, bytes));
  FB_CUDA(cudaMalloc(&scratch, bytes));
  FB_CUDA(cudaMemcpyAsync(x, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
  dom.transform(x, scratch, kind, ctx->stream);
  FB_CUDA(cudaMemcpyAsync(data, x, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  FB_CUDA(cudaGetLastError());
  cudaFree(x); cudaFree(scratch);
  dom.destroy();
  return FB_OK;
} FB_ABI_CATCH_INT

int fb_test_h(fb_ctx* ctx_, int log_n, const uint64_t* a, const uint64_t* b, const uint64_t* c,
              uint64_t* out, float* ms) try {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  if (!ctx || !a || !b || !c) return FB_ERR_ARG;
  FB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  NttDomain dom;
  if (dom.init(log_n, st) != 0) { set_error("domain init failed"); return FB_ERR_DOMAIN; }
  const size_t bytes = sizeof(Fr) << log_n;
  Fr *ev[3], *scratch;
  const uint64_t* src[3] = {a, b, c};
  for (int i = 0; i < 3; i++) FB_CUDA(cudaMalloc(&ev[i], bytes));
  FB_CUDA(cudaMalloc(&scratch, bytes));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  const int reps = ms ? 3 : 1;
  for (int rep = 0; rep < reps; rep++) {
    for (int i = 0; i < 3; i++) FB_CUDA(cudaMemcpyAsync(ev[i], src[i], bytes, cudaMemcpyHostToDevice, st));
    cudaEventRecord(e0, st);
    for (int i = 0; i < 3; i++) dom.ifft_then_coset_fft(ev[i], st);
    dom.pointwise_then_icoset_fft(ev[0], ev[1], ev[2], st);
    cudaEventRecord(e1, st);
    FB_CUDA(cudaStreamSynchronize(st));
    float t;
    cudaEventElapsedTime(&t, e0, e1);
    best = std::min(best, t);
  }
  if (ms) *ms = best;
  if (out) {
    dom.bitrev(scratch, ev[0], st);
    FB_CUDA(cudaMemcpyAsync(out, scratch, bytes - sizeof(Fr), cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaStreamSynchronize(st));
  }
  FB_CUDA(cudaGetLastError());
  for (int i = 0; i < 3; i++) cudaFree(ev[i]);
  cudaFree(scratch);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  dom.destroy();
  return FB_OK;
} FB_ABI_CATCH_INT

int fb_test_msm(fb_ctx* ctx_, int group, const uint8_t* bases_raw, const uint64_t* scalars,
                uint64_t n, uint8_t* result_raw, int reps, float* ms_per_rep) try {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  if (!ctx || !bases_raw || !scalars || !result_raw || (group != 1 && group != 2) || n == 0 ||
      n >= (1ull << 31))
    return FB_ERR_ARG;
  FB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t psz = group == 1 ? 64 : 128;
  void *dbases, *dres;
  Fr* dsc;
  MsmPlan plan = MsmPlan::make((uint32_t)n, g_msm_tables == 1);  // fb_set_msm_tables(1) selects window tables
  FB_CUDA(cudaMalloc(&dbases, plan.table_points() * psz));
  FB_CUDA(cudaMalloc(&dsc, n * 32));
  FB_CUDA(cudaMalloc(&dres, sizeof(G2XYZZ) * MSM_VBITS));
  FB_CUDA(cudaMemcpyAsync(dbases, bases_raw, n * psz, cudaMemcpyHostToDevice, st));
  FB_CUDA(cudaMemcpyAsync(dsc, scalars, n * 32, cudaMemcpyHostToDevice, st));
  if ((group == 1 ? msm_build_table_g1((G1Affine*)dbases, plan, st) : msm_build_table_g2((G2Affine*)dbases, plan, st)) != 0) {
    set_error("msm window table build failed");
    return FB_ERR_CUDA;
  }
  MsmScratch scr;
  if (scr.alloc(&plan, 1, group == 2) != 0) { set_error("msm scratch alloc failed"); return FB_ERR_CUDA; }
  std::vector<G2XYZZ> hres(MSM_VBITS);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  if (reps < 1) reps = 1;
  float best = 1e30f;
  int rc = 0;
  const int nbits = plan.vbits();
  for (int rep = 0; rep < reps && !rc; rep++) {
    // timed: digits + sort + accumulate + reduce on the device, then the Horner tail on the host
    auto h0 = std::chrono::steady_clock::now();
    cudaEventRecord(e0, st);
    if (group == 1) rc = msm_g1((const G1Affine*)dbases, dsc, nullptr, plan, scr, (G1XYZZ*)dres, false, st);
    else rc = msm_g2((const G2Affine*)dbases, dsc, nullptr, plan, scr, (G2XYZZ*)dres, false, st);
    FB_CUDA(cudaMemcpyAsync(hres.data(), dres, (group == 1 ? sizeof(G1XYZZ) : sizeof(G2XYZZ)) * MSM_VBITS,
                            cudaMemcpyDeviceToHost, st));
    cudaEventRecord(e1, st);
    FB_CUDA(cudaStreamSynchronize(st));
    if (group == 1) {
      Affine<HFq> a = to_affine(msm_horner_host<HFq>((const G1XYZZ*)hres.data(), nbits));
      G1Affine o{a.x.to(), a.y.to()};
      memcpy(result_raw, &o, 64);
    } else {
      Affine<HFq2> a = to_affine(msm_horner_host<HFq2>(hres.data(), nbits));
      G2Affine o{a.x.to(), a.y.to()};
      memcpy(result_raw, &o, 128);
    }
    auto h1 = std::chrono::steady_clock::now();
    float t = (float)std::chrono::duration<double, std::milli>(h1 - h0).count();
    best = std::min(best, t);
  }
  if (rc) { set_error("msm failed %d", rc); return FB_ERR_CUDA; }
  if (ms_per_rep) *ms_per_rep = best;
  FB_CUDA(cudaGetLastError());
  cudaFree(dbases); cudaFree(dsc); cudaFree(dres);
  scr.release();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return FB_OK;
} FB_ABI_CATCH_INT

int fb_probe_imad(fb_ctx* ctx_, double* mac_per_s) try {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  if (!ctx || !mac_per_s) return FB_ERR_ARG;
  FB_CUDA(cudaSetDevice(ctx->device));
  uint64_t* d;
  FB_CUDA(cudaMalloc(&d, 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 4096, blocks = 148 * 8, threads = 256;
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0, ctx->stream);
    k_probe_imad<<<blocks, threads, 0, ctx->stream>>>(d, 12345u + rep, iters);
    cudaEventRecord(e1, ctx->stream);
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double rate = (double)blocks * threads * iters * 16 / (ms * 1e-3);
    if (rep > 0) best = std::max(best, rate);
  }
  *mac_per_s = best;
  cudaFree(d);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return FB_OK;
} FB_ABI_CATCH_INT

// which: 0 carry-chained wide MAC rate (MAC/s), 1 mul_ptx rate, 2 mul_c rate (mul/s);
// threads per block and blocks per SM selectable to see the occupancy dependence
int fb_probe_rate(fb_ctx* ctx_, int which, int threads, int blocks_per_sm, double* per_s) try {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  if (!ctx || !per_s) return FB_ERR_ARG;
  FB_CUDA(cudaSetDevice(ctx->device));
  void* d;
  FB_CUDA(cudaMalloc(&d, 64));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = 148 * blocks_per_sm;
  const int iters = which == 0 ? 1024 : 512;
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0, ctx->stream);
    if (which == 0) k_probe_madc<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)d, 777u + rep, iters);
    else if (which == 1) k_probe_fr_mul_v<0><<<blocks, threads, 0, ctx->stream>>>((Fr*)d, iters);
    else if (which == 9) k_probe_imad3<<<blocks, threads, 0, ctx->stream>>>((uint64_t*)d, 777u + rep, iters);
    else if (which == 10) k_probe_imad_chain<<<blocks, threads, 0, ctx->stream>>>((uint64_t*)d, 777u + rep, iters);
    else if (which == 5) k_probe_mac_shape<0><<<blocks, threads, 0, ctx->stream>>>((uint64_t*)d, 777u + rep, iters);
    else if (which == 6) k_probe_mac_shape<1><<<blocks, threads, 0, ctx->stream>>>((uint64_t*)d, 777u + rep, iters);
    else if (which == 7) k_probe_mac_shape<2><<<blocks, threads, 0, ctx->stream>>>((uint64_t*)d, 777u + rep, iters);
    else if (which == 8) k_probe_mac_shape<3><<<blocks, threads, 0, ctx->stream>>>((uint64_t*)d, 777u + rep, iters);
    else k_probe_fr_mul_v<1><<<blocks, threads, 0, ctx->stream>>>((Fr*)d, iters);
    cudaEventRecord(e1, ctx->stream);
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double work = which == 0 ? (double)iters * 4 * 4 : (which >= 5 ? (double)iters * 8 : (double)iters * 2);
    double rate = (double)blocks * threads * work / (ms * 1e-3);
    if (rep > 0) best = std::max(best, rate);
  }
  *per_s = best;
  cudaFree(d);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return FB_OK;
} FB_ABI_CATCH_INT

// Host only: the Pippenger plan MsmPlan::make picks for n points (window bits, digits per scalar, log2 entries per
// accumulation task) with or without window tables -- lets the CPU tests pin the plans of the benchmark sizes.
int fb_test_msm_plan(uint32_t n, int table, int* c, int* W, int* task_log) try {
  const MsmPlan p = MsmPlan::make(n, table != 0);
  if (c) *c = p.c;
  if (W) *W = p.W;
  if (task_log) *task_log = p.task_log;
  return FB_OK;
} FB_ABI_CATCH_INT

int fb_probe_fr_mul(fb_ctx* ctx_, double* mul_per_s) try {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  if (!ctx || !mul_per_s) return FB_ERR_ARG;
  FB_CUDA(cudaSetDevice(ctx->device));
  Fr* d;
  FB_CUDA(cudaMalloc(&d, 32));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 512, blocks = 148 * 8, threads = 256;
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0, ctx->stream);
    k_probe_fr_mul<<<blocks, threads, 0, ctx->stream>>>(d, iters);
    cudaEventRecord(e1, ctx->stream);
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double rate = (double)blocks * threads * iters * 2 / (ms * 1e-3);
    if (rep > 0) best = std::max(best, rate);
  }
  *mul_per_s = best;
  cudaFree(d);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return FB_OK;
} FB_ABI_CATCH_INT

}  // extern "C"
