// Internal structures shared by the translation units of libfawkes_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <exception>
#include <string>
#include <vector>

#include "../../include/fawkes_b200.h"
#include "ec.cuh"
#include "host_fq.h"
#include "msm.cuh"
#include "ntt.cuh"

namespace fb {

void set_error(const char* fmt, ...);
// MSM window tables (msm.cuh): -1 auto (on when the tables fit in free HBM with headroom), 0 off, 1 on
extern int g_msm_tables;

// ---- instrumentation (bench.py: gpu_launches and the live roofline timing) -------------
extern std::atomic<unsigned long long> g_launches;  // kernels launched by this library
inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
enum { KSTAT_ACC_G1 = 0, KSTAT_ACC_G2 = 1, KSTAT_NTT = 2, KSTAT_EXCHANGE = 3, KSTAT_SORT = 4, KSTAT_REDUCE = 5, KSTAT_KINDS = 6 };
void kstat_begin(int kind, cudaStream_t st);    // no-ops unless enabled
void kstat_end(int kind, cudaStream_t st);
#define FB_CUDA(call)                                                              \
  do {                                                                             \
    cudaError_t _e = (call);                                                       \
    if (_e != cudaSuccess) {                                                       \
      fb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return FB_ERR_CUDA;                                                      \
    }                                                                              \
  } while (0)

// Every entry point of the C ABI is a function-try-block closed by one of these: a C++ exception of the host code
// (std::bad_alloc while parsing a multi-GB gate stream, std::system_error from thread creation) must not unwind
// into the caller's frames -- for the Rust shim that is undefined behaviour.  The caller sees FB_ERR_HOST.
#define FB_ABI_CATCH_INT                                                   \
  catch (const std::exception& e) {                                        \
    fb::set_error("host exception: %s", e.what());                         \
    return FB_ERR_HOST;                                                    \
  }                                                                        \
  catch (...) {                                                            \
    fb::set_error("unknown host exception");                               \
    return FB_ERR_HOST;                                                    \
  }
#define FB_ABI_CATCH_VOID \
  catch (...) {           \
  }

// R1CS in CSR form over the concatenated variable vector w = [inputs | aux].
struct HostCsr {
  std::vector<uint32_t> rowptr[3];  // n_gates + 1 each
  std::vector<uint32_t> col[3];
  std::vector<uint32_t> cidx[3];    // 0 -> ONE, 1 -> -ONE, k>=2 -> coef[k-2]
  std::vector<Fr> coef;             // Montgomery
  uint32_t n_gates = 0;
};

struct DevCsr {
  uint32_t* rowptr[3] = {nullptr, nullptr, nullptr};
  uint32_t* col[3] = {nullptr, nullptr, nullptr};
  uint32_t* cidx[3] = {nullptr, nullptr, nullptr};
  Fr* coef = nullptr;
  uint32_t n_gates = 0;
  uint64_t nnz[3] = {0, 0, 0};
};

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;   // R1CS eval, H pipeline, H MSM, copies
  cudaStream_t aux[3] = {nullptr, nullptr, nullptr};  // L / A / B MSMs run beside the H pipeline
  cudaEvent_t aux_done[3] = {nullptr, nullptr, nullptr};
  // multi-GPU (fb_dist_init): rank / world of this process and the NTT exchange transport
  int rank = 0, world = 1;
  void* exchange = nullptr;  // NttExchange*
  // staged upload of pageable witness buffers (api.cu: upload_host): pinned ring + copy streams, made on first use
  void* stage = nullptr;
  cudaStream_t copy_stream[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t stage_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t copy_done[4] = {nullptr, nullptr, nullptr, nullptr};
};

struct ProvingKey {
  Ctx* ctx = nullptr;
  uint32_t n_in = 0, n_aux = 0, n_rows = 0;
  int k = 0;           // log2 domain size
  uint64_t m = 0;
  // verifying-key points needed by the prover (host, Montgomery affine)
  G1Affine alpha_g1, beta_g1, delta_g1;
  G2Affine beta_g2, delta_g2;
  // base arrays in HBM
  G1Affine* h = nullptr;   // m-1 points, bit-reversed order
  G1Affine* l = nullptr;   // n_aux
  G1Affine* a = nullptr;   // len_a
  G1Affine* b1 = nullptr;  // len_b
  G2Affine* b2 = nullptr;  // len_b
  uint32_t len_h = 0, len_a = 0, len_b = 0;
  uint32_t* a_map = nullptr;  // into w
  uint32_t* b_map = nullptr;
  DevCsr csr;
  NttDomain dom;
  // per-prove workspace
  Fr* w = nullptr;         // n_in + n_aux
  Fr* ev[3] = {nullptr, nullptr, nullptr};  // m each
  Fr* scratch = nullptr;   // m (h_out permutation)
  MsmScratch msm[4];       // H, L, A, B (B_g1 and B_g2 share one sort)
  MsmPlan plan_h, plan_l, plan_a, plan_b;
  void* results = nullptr;       // 5 x MSM_VBITS G2XYZZ-sized slots on device: bit sums of H L A B1 B2
  cudaEvent_t msm_done[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  void* results_host = nullptr;  // pinned
  // base-index shard handled by this key (multi-GPU): fractions [shard, shard+1)/nshards
  int shard = 0, nshards = 1;
  // distributed H pipeline (0 = evaluation + H replicated on every rank): R1CS rows are dealt
  // cyclically over 2^dist_g ranks, ev[] are local arrays of m >> dist_g, xtmp the exchange buffers
  int dist_g = 0;
  uint32_t n_gates_global = 0;
  Fr* xtmp[3] = {nullptr, nullptr, nullptr};
  void* host_tables = nullptr;  // fixed-base tables of delta/alpha/beta for the host-side assembly
  uint64_t table_bytes = 0;     // HBM held by the MSM window tables beyond the plain base arrays
  // fb_prove_batch: extra per-prove workspaces ("slots") that share this key's immutable arrays (bases and
  // window tables, CSR, twiddles, host tables) and own everything a prove writes, each with its own streams
  bool is_slot = false;
  std::vector<ProvingKey*> slots;
  void* batch = nullptr;  // BatchWork* (api.cu): workspaces of the batched path of fb_prove_batch
  // small keys: the whole device side of a prove (upload, ~70 kernels on four streams, result copies)
  // captured once as a CUDA graph and replayed per proof (api.cu: prove_graph)
  cudaGraphExec_t graph_exec = nullptr;
  void* w_stage = nullptr;               // pinned staging copy of the witness (fixed address for the graph)
  unsigned long long graph_launches = 0; // kernels inside the graph (for fb_launch_count)
  bool graph_failed = false;             // capture was refused once: stay on the stream path
};

struct Circuit {
  HostCsr csr;
  uint32_t n_in = 0, n_aux = 0;
  std::vector<Fr> inputs, aux;  // synthetic circuits only (Montgomery)
};

struct ParamsView {  // offsets into bellman Parameters bytes
  const uint8_t *alpha_g1, *beta_g1, *beta_g2, *gamma_g2, *delta_g1, *delta_g2;
  const uint8_t *ic, *h, *l, *a, *b1, *b2;
  uint32_t n_ic, n_h, n_l, n_a, n_b1, n_b2;
};
int parse_params(const uint8_t* b, size_t len, ParamsView& v);  // api.cu
int key_dist_g(const Ctx* ctx, int k, int shard, int nshards);    // api.cu: how a key on this context shards H
void key_h_range(uint64_t m, int dist_g, int shard, int nshards, uint64_t* lo, uint64_t* cnt);

// pk.cu
int parse_gates_to_csr(const uint8_t* raw, size_t len, uint32_t n_in, uint32_t n_aux, HostCsr& out);
int brotli_decode(const uint8_t* in, size_t len, std::vector<uint8_t>& out, size_t max_out);
int upload_csr(const HostCsr& h, DevCsr& d, cudaStream_t st);
// ingest.cu: the same CSR with the per-term work (parse, range checks, Montgomery form, dictionary) on the GPU
struct Ctx;
int parse_gates_device(Ctx* ctx, const uint8_t* raw, size_t len, uint32_t n_in, uint32_t n_aux, HostCsr& out,
                       float* times_ms);
void free_csr(DevCsr& d);
// flags: FB_LOAD_CHECKED | FB_LOAD_NO_INFINITY (include/fawkes_b200.h)
int decode_g1_be(const uint8_t* host_be, uint64_t n, G1Affine* dev_out, int flags, cudaStream_t st);
int decode_g2_be(const uint8_t* host_be, uint64_t n, G2Affine* dev_out, int flags, cudaStream_t st);
int host_decode_g1(const uint8_t* be, G1Affine& out, int extra_flags = 0);  // always checked; 0 or a point_error code
int host_decode_g2(const uint8_t* be, G2Affine& out, int extra_flags = 0);
const char* point_error(int code);
void host_encode_g1(const G1Affine& p, uint8_t* be);
void host_encode_g2(const G2Affine& p, uint8_t* be);

// dist.cu
struct NttExchange;
NttExchange* dist_exchange(Ctx* ctx);
int dist_all_gather_inplace(Ctx* ctx, void* buf, size_t chunk_bytes, cudaStream_t st);
int dist_gather_partials(Ctx* ctx, const uint8_t* mine, uint8_t* all, cudaStream_t st);  // 640 B per rank
void dist_destroy(Ctx* ctx);  // communicator + buffers (fb_shutdown)

// prove.cu
int eval_r1cs(const DevCsr& csr, const Fr* w, uint32_t n_in, Fr* a, Fr* b, Fr* c, uint64_t m,
              cudaStream_t st);
int eval_r1cs_batch(const DevCsr& csr, const Fr* w, uint64_t wstride, uint32_t n_in, Fr* a, Fr* b, Fr* c, uint64_t m,
                    uint32_t count, cudaStream_t st);
int eval_r1cs_cyclic(const DevCsr& local_csr, const Fr* w, uint32_t n_in, uint32_t n_gates_global, int g, int rank,
                     Fr* a, Fr* b, Fr* c, uint64_t ml, cudaStream_t st);

}  // namespace fb
