// Multi-GPU plumbing of the distributed H pipeline: the all-to-all between the cyclic and block
// layouts of ntt.cu, over NCCL (NVLink / NVSwitch) with one process per GPU.
//
// The reference has no distributed path at all (single process); this is
// north_star subsystem (3): "the four-step transpose sharded across GPUs via NCCL all-to-all".
// libnccl is bound at run time (dlopen) so the library still loads on a machine without it; the
// host application creates the unique id on rank 0, ships it to the other ranks by any means
// (bench.py: torch.distributed broadcast) and every rank calls fb_dist_init.
#include <dlfcn.h>

#include <condition_variable>
#include <mutex>
#include <thread>

#include "internal.h"

namespace fb {

// ---- the slice of the NCCL API we use (types restated from nccl.h) ----
typedef void* ncclComm_t;
struct ncclUniqueId_ { char internal[128]; };
static const int kNcclUint8 = 1;  // ncclDataType_t::ncclUint8
struct NcclApi {
  int (*GetUniqueId)(ncclUniqueId_*);
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId_, int);
  int (*CommDestroy)(ncclComm_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(int);
  bool ok = false;
};
static NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (lib) {
      api.GetUniqueId = (int (*)(ncclUniqueId_*))dlsym(lib, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId_, int))dlsym(lib, "ncclCommInitRank");
      api.CommDestroy = (int (*)(ncclComm_t))dlsym(lib, "ncclCommDestroy");
      api.GroupStart = (int (*)())dlsym(lib, "ncclGroupStart");
      api.GroupEnd = (int (*)())dlsym(lib, "ncclGroupEnd");
      api.Send = (int (*)(const void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(lib, "ncclSend");
      api.Recv = (int (*)(void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(lib, "ncclRecv");
      api.AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(lib, "ncclAllGather");
      api.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
      api.ok = api.GetUniqueId && api.CommInitRank && api.GroupStart && api.GroupEnd && api.Send && api.Recv;
    }
  });
  return api;
}

struct NcclExchange : NttExchange {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  // combine step of a sharded prove: world x 640-byte partial sums (device + pinned host mirror)
  uint8_t* gather_dev = nullptr;
  uint8_t* gather_host = nullptr;
  // exchanges of one array overlap the transforms of the next (ntt.cu: dist_h_pipeline)
  cudaStream_t side = nullptr;
  cudaEvent_t evs[12] = {};
  cudaStream_t side_stream() override { return side; }
  cudaEvent_t event(int i) override { return evs[i]; }
  int all_to_all(const Fr* const* send, Fr* const* recv, int narrays, uint64_t count, cudaStream_t st) override {
    NcclApi& n = nccl();
    int rc = n.GroupStart();
    const size_t bytes = count * sizeof(Fr);
    for (int v = 0; v < narrays && !rc; v++)
      for (int r = 0; r < world && !rc; r++) {
        rc = n.Send(send[v] + (uint64_t)r * count, bytes, kNcclUint8, r, comm, st);
        if (!rc) rc = n.Recv(recv[v] + (uint64_t)r * count, bytes, kNcclUint8, r, comm, st);
      }
    int rc2 = n.GroupEnd();
    if (rc || rc2) {
      set_error("NCCL all-to-all failed: %s", n.GetErrorString ? n.GetErrorString(rc ? rc : rc2) : "?");
      return FB_ERR_CUDA;
    }
    count_launch(1);
    return 0;
  }
};

NttExchange* dist_exchange(Ctx* ctx) { return reinterpret_cast<NttExchange*>(ctx->exchange); }

// In-place all-gather of equal chunks over NVLink: rank r has filled buf[r*chunk_bytes ..) already.
int dist_all_gather_inplace(Ctx* ctx, void* buf, size_t chunk_bytes, cudaStream_t st) {
  NcclExchange* x = reinterpret_cast<NcclExchange*>(ctx->exchange);
  NcclApi& n = nccl();
  if (!x || !n.AllGather) { set_error("NCCL all-gather unavailable"); return FB_ERR_CUDA; }
  int rc = n.AllGather((const uint8_t*)buf + (size_t)x->rank * chunk_bytes, buf, chunk_bytes, kNcclUint8, x->comm, st);
  if (rc) { set_error("ncclAllGather: %s", n.GetErrorString ? n.GetErrorString(rc) : "?"); return FB_ERR_CUDA; }
  count_launch(1);
  return 0;
}

// Combine step of a sharded prove (north_star (4): "partial sums combined over NVLink"): every rank contributes
// its five affine partial sums (640 B, the fb_prove_partial layout) and receives everybody's, in rank order.
// One ncclAllGather over the library's communicator on the prove stream; `all` is host memory, world x 640 B.
int dist_gather_partials(Ctx* ctx, const uint8_t* mine, uint8_t* all, cudaStream_t st) {
  NcclExchange* x = reinterpret_cast<NcclExchange*>(ctx->exchange);
  NcclApi& n = nccl();
  if (!x || !n.AllGather) { set_error("NCCL all-gather unavailable"); return FB_ERR_CUDA; }
  const size_t total = (size_t)x->world * 640;
  if (!x->gather_dev) {
    FB_CUDA(cudaMalloc(&x->gather_dev, total));
    FB_CUDA(cudaMallocHost(&x->gather_host, total));
  }
  memcpy(x->gather_host + (size_t)x->rank * 640, mine, 640);
  FB_CUDA(cudaMemcpyAsync(x->gather_dev + (size_t)x->rank * 640, x->gather_host + (size_t)x->rank * 640, 640,
                          cudaMemcpyHostToDevice, st));
  int rc = n.AllGather(x->gather_dev + (size_t)x->rank * 640, x->gather_dev, 640, kNcclUint8, x->comm, st);
  if (rc) { set_error("ncclAllGather: %s", n.GetErrorString ? n.GetErrorString(rc) : "?"); return FB_ERR_CUDA; }
  count_launch(1);
  FB_CUDA(cudaMemcpyAsync(x->gather_host, x->gather_dev, total, cudaMemcpyDeviceToHost, st));
  FB_CUDA(cudaStreamSynchronize(st));
  memcpy(all, x->gather_host, total);
  return FB_OK;
}

void dist_destroy(Ctx* ctx) {
  NcclExchange* x = reinterpret_cast<NcclExchange*>(ctx->exchange);
  if (!x) return;
  if (x->gather_dev) cudaFree(x->gather_dev);
  if (x->gather_host) cudaFreeHost(x->gather_host);
  if (x->side) cudaStreamDestroy(x->side);
  for (auto& e : x->evs) if (e) cudaEventDestroy(e);
  NcclApi& n = nccl();
  if (x->comm && n.CommDestroy) n.CommDestroy(x->comm);
  delete x;
  ctx->exchange = nullptr;
  ctx->world = 1;
  ctx->rank = 0;
}

// ---- single-process stand-in used by the one-GPU test: G host threads, one per virtual rank ----
struct LocalWorld {
  int world;
  std::mutex mu;
  std::condition_variable cv;
  int waiting = 0;
  uint64_t generation = 0;
  std::vector<const Fr* const*> send;
  explicit LocalWorld(int w) : world(w), send(w) {}
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    uint64_t gen = generation;
    if (++waiting == world) { waiting = 0; generation++; cv.notify_all(); }
    else cv.wait(lk, [&] { return generation != gen; });
  }
};
struct LocalExchange : NttExchange {
  LocalWorld* w;
  int rank;
  cudaStream_t side = nullptr;  // set: dist_h_pipeline takes its overlapped schedule (exchanges on this stream)
  cudaEvent_t evs[12] = {};
  cudaStream_t side_stream() override { return side; }
  cudaEvent_t event(int i) override { return evs[i]; }
  int all_to_all(const Fr* const* send, Fr* const* recv, int narrays, uint64_t count, cudaStream_t st) override {
    cudaStreamSynchronize(st);          // my send buffers are final
    w->send[rank] = send;
    w->barrier();                       // everybody's are
    for (int v = 0; v < narrays; v++)
      for (int r = 0; r < w->world; r++)
        cudaMemcpyAsync(recv[v] + (uint64_t)r * count, w->send[r][v] + (uint64_t)rank * count, count * sizeof(Fr),
                        cudaMemcpyDeviceToDevice, st);
    cudaStreamSynchronize(st);
    w->barrier();                       // nobody overwrites a send buffer that is still being read
    return 0;
  }
};

}  // namespace fb

using namespace fb;

extern "C" {

int fb_dist_unique_id(uint8_t id[128]) try {
  if (!id) return FB_ERR_ARG;
  NcclApi& n = nccl();
  if (!n.ok) { set_error("libnccl.so.2 not available"); return FB_ERR_CUDA; }
  ncclUniqueId_ u;
  int rc = n.GetUniqueId(&u);
  if (rc) { set_error("ncclGetUniqueId: %s", n.GetErrorString(rc)); return FB_ERR_CUDA; }
  memcpy(id, u.internal, 128);
  return FB_OK;
} FB_ABI_CATCH_INT

int fb_dist_init(fb_ctx* ctx_, int rank, int world, const uint8_t id[128]) try {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  if (!ctx || !id || world < 1 || rank < 0 || rank >= world) { set_error("fb_dist_init: bad argument"); return FB_ERR_ARG; }
  NcclApi& n = nccl();
  if (!n.ok) { set_error("libnccl.so.2 not available"); return FB_ERR_CUDA; }
  FB_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId_ u;
  memcpy(u.internal, id, 128);
  NcclExchange* x = new NcclExchange();
  x->rank = rank;
  x->world = world;
  int rc = n.CommInitRank(&x->comm, world, u, rank);
  if (rc) { set_error("ncclCommInitRank: %s", n.GetErrorString(rc)); delete x; return FB_ERR_CUDA; }
  if (!getenv("FB_DIST_NO_OVERLAP")) {  // without the side stream the exchanges run on the compute stream
    bool ok = cudaStreamCreateWithFlags(&x->side, cudaStreamNonBlocking) == cudaSuccess;
    for (auto& e : x->evs) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { cudaGetLastError(); if (x->side) cudaStreamDestroy(x->side); x->side = nullptr; }
  }
  if (ctx->exchange) dist_destroy(ctx);
  ctx->exchange = x;
  ctx->rank = rank;
  ctx->world = world;
  return FB_OK;
} FB_ABI_CATCH_INT

// One-GPU check of the distributed H pipeline: 2^g virtual ranks (host threads, own streams) with the
// exchange done by device-to-device copies.  a, b, c: row evaluations [2^log_n][4]; out: H
// coefficients [2^log_n - 1][4] in natural order.
int fb_test_dist_h(fb_ctx* ctx_, int log_n, int g, const uint64_t* a, const uint64_t* b, const uint64_t* c,
                   uint64_t* out) try {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  if (!ctx || !a || !b || !c || !out) return FB_ERR_ARG;
  FB_CUDA(cudaSetDevice(ctx->device));
  NttDomain dom;
  if (dom.init(log_n, ctx->stream) != 0) { set_error("domain init failed"); return FB_ERR_DOMAIN; }
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (!dom.dist_supported(g)) { dom.destroy(); set_error("2^%d over 2^%d ranks is not supported", log_n, g); return FB_ERR_ARG; }
  const int G = 1 << g, kl = log_n - g;
  const uint64_t m = 1ull << log_n, ml = 1ull << kl;
  std::vector<Fr> result(m), result_ov(m);
  const uint64_t* src[3] = {a, b, c};
  // both schedules of dist_h_pipeline: exchanges on the compute stream, then on a side stream under the transforms
  for (int overlap = 0; overlap < 2; overlap++) {
    LocalWorld world(G);
    std::vector<std::thread> th;
    std::vector<int> rcs(G, 0);
    std::vector<Fr>& res = overlap ? result_ov : result;
    for (int r = 0; r < G; r++) {
      th.emplace_back([&, r] {
        cudaSetDevice(ctx->device);
        cudaStream_t st;
        cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        Fr *ev[3], *tmp[3];
        std::vector<Fr> host(ml);
        for (int v = 0; v < 3; v++) {
          cudaMalloc(&ev[v], ml * sizeof(Fr));
          cudaMalloc(&tmp[v], ml * sizeof(Fr));
          for (uint64_t j = 0; j < ml; j++) memcpy(&host[j], src[v] + 4 * ((j << g) | (uint64_t)r), 32);  // cyclic
          cudaMemcpyAsync(ev[v], host.data(), ml * sizeof(Fr), cudaMemcpyHostToDevice, st);
          cudaStreamSynchronize(st);
        }
        LocalExchange x;
        x.w = &world;
        x.rank = r;
        if (overlap) {
          cudaStreamCreateWithFlags(&x.side, cudaStreamNonBlocking);
          for (auto& e : x.evs) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        }
        rcs[r] = dom.dist_h_pipeline(ev, tmp, g, r, &x, st);
        cudaStreamSynchronize(st);
        if (cudaGetLastError() != cudaSuccess) rcs[r] = -9;
        // block layout, bit-reversed coefficient order: global position r * ml + j
        cudaMemcpy(res.data() + (uint64_t)r * ml, ev[0], ml * sizeof(Fr), cudaMemcpyDeviceToHost);
        for (int v = 0; v < 3; v++) { cudaFree(ev[v]); cudaFree(tmp[v]); }
        if (x.side) cudaStreamDestroy(x.side);
        for (auto& e : x.evs) if (e) cudaEventDestroy(e);
        cudaStreamDestroy(st);
      });
    }
    for (auto& t : th) t.join();
    for (int r = 0; r < G; r++)
      if (rcs[r]) { dom.destroy(); set_error("virtual rank %d failed (%d), overlap=%d", r, rcs[r], overlap); return FB_ERR_CUDA; }
  }
  dom.destroy();
  if (memcmp(result.data(), result_ov.data(), m * sizeof(Fr)) != 0) {
    set_error("the overlapped exchange schedule gives a different H");
    return FB_ERR_CUDA;
  }
  for (uint64_t p = 0; p < m; p++) {
    uint64_t i = 0;
    for (int bit = 0; bit < log_n; bit++) i |= ((p >> bit) & 1) << (log_n - 1 - bit);
    if (i < m - 1) memcpy(out + 4 * i, &result[p], 32);
  }
  return FB_OK;
} FB_ABI_CATCH_INT

}  // extern "C"
