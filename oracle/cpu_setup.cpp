// CPU restatement of the harness side of the reference's Groth16 path -- C++17, threads:
//   * the synthetic random R1CS + witness of SURVEY.md section 8(d) (same generator as oracle/synth.py),
//   * `setup` with an explicit trapdoor (bellman_ce generator; call site
//     fawkes-crypto/src/backend/bellman_groth16/setup.rs:17-20), output in bellman's Parameters bytes.
//
// TEST INFRASTRUCTURE ONLY (same rule as cpu_prover.cpp).  With these two the CPU arm of bench.py
// (`--impl reference`, `cpu_baseline`) builds its circuit and its keys without touching the product library,
// at the full benchmark sizes (2^24 rows), so the proofs of the two arms can be compared byte for byte.
//
// What each part follows:
//   inputize row / CNum product rows     fawkes-crypto/src/circuit/r1cs/cs.rs:309-318, num.rs:253-272
//   row set incl. `input_i * 0 = 0`      fawkes-crypto/src/backend/bellman_groth16/mod.rs:61-102
//   Lagrange values by ifft of powers, a/b/l/ic/h queries, infinity filtering   bellman_ce generator [App. C.4]
//   generators                           (1, 2) and the EIP-197 G2 generator (harness convention)
#include "cpu_common.h"

#include <cstdlib>

namespace {

struct SplitMix64 {
  u64 s;
  explicit SplitMix64(u64 seed) : s(seed) {}
  u64 next() {
    s += 0x9E3779B97F4A7C15ull;
    u64 z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  Fr fr_mont() {  // 4 words, top limb masked to 62 bits, reject >= r; then to Montgomery form
    for (;;) {
      Fr h;
      for (int i = 0; i < 4; i++) h.v[i] = next();
      h.v[3] &= (1ull << 62) - 1;
      if (!Fr::geq(h.v)) return h.to_mont();
    }
  }
};

struct OCircuit {
  uint32_t n_in = 0, n_aux = 0, n_gates = 0;
  std::vector<uint32_t> rowptr[3], col[3];
  std::vector<Fr> coef[3];  // one Montgomery coefficient per term
  std::vector<Fr> inputs, aux;
};

constexpr int N_INIT_AUX = 16;

// ---- fixed-base scalar multiplication: T[w][d-1] = d * 2^(WB w) * G, affine ---------------------------
constexpr int WB = 14;
constexpr int NWIN = (254 + WB - 1) / WB;
constexpr size_t WSZ = (size_t(1) << WB) - 1;

template <class F>
static void batch_to_affine(const Jac<F>* in, size_t n, Aff<F>* out) {
  std::vector<F> pre(n);
  F acc = F::one();
  for (size_t i = 0; i < n; i++) {
    pre[i] = acc;
    if (!in[i].is_zero()) acc = acc * in[i].z;
  }
  F iv = acc.inv();
  for (size_t i = n; i-- > 0;) {
    if (in[i].is_zero()) { out[i] = {F::zero(), F::zero(), true}; continue; }
    F zi = iv * pre[i];
    iv = iv * in[i].z;
    F zi2 = zi.sqr();
    out[i] = {in[i].x * zi2, in[i].y * zi2 * zi, false};
  }
}

template <class F>
struct FixedBaseTable {
  std::vector<Aff<F>> t;  // NWIN * WSZ
  void build(const Aff<F>& g, int nthreads) {
    t.resize((size_t)NWIN * WSZ);
    std::vector<Jac<F>> base(NWIN);
    Jac<F> b = Jac<F>::zero();
    b.add_mixed(g);
    for (int w = 0; w < NWIN; w++) {
      base[w] = b;
      for (int i = 0; i < WB; i++) b.dbl();
    }
    std::vector<std::function<void()>> tasks;
    for (int w = 0; w < NWIN; w++)
      tasks.push_back([this, w, &base] {
        std::vector<Jac<F>> row(WSZ);
        Jac<F> acc = Jac<F>::zero();
        for (size_t d = 0; d < WSZ; d++) {
          acc.add(base[w]);
          row[d] = acc;
        }
        batch_to_affine(row.data(), WSZ, t.data() + (size_t)w * WSZ);
      });
    run_tasks(nthreads, tasks);
  }
  Jac<F> mul(const Fr& s_mont) const {
    const Fr s = s_mont.from_mont();
    Jac<F> acc = Jac<F>::zero();
    for (int w = 0; w < NWIN; w++) {
      const int pos = w * WB, limb = pos >> 6, sh = pos & 63;
      u64 d = s.v[limb] >> sh;
      if (sh + WB > 64 && limb + 1 < 4) d |= s.v[limb + 1] << (64 - sh);
      d &= WSZ;
      if (d) acc.add_mixed(t[(size_t)w * WSZ + d - 1]);
    }
    return acc;
  }
};

static void fq_to_be(const Fq& m, uint8_t* be) {
  const Fq c = m.from_mont();
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 8; j++) be[(3 - i) * 8 + j] = (uint8_t)(c.v[i] >> (56 - 8 * j));
}
static void encode(const Aff<Fq>& p, uint8_t* o) {
  if (p.inf) { memset(o, 0, 64); o[0] = 0x40; return; }
  fq_to_be(p.x, o);
  fq_to_be(p.y, o + 32);
}
static void encode(const Aff<Fq2>& p, uint8_t* o) {
  if (p.inf) { memset(o, 0, 128); o[0] = 0x40; return; }
  fq_to_be(p.x.c1, o);
  fq_to_be(p.x.c0, o + 32);
  fq_to_be(p.y.c1, o + 64);
  fq_to_be(p.y.c0, o + 96);
}

template <class F>
static void emit_points(const FixedBaseTable<F>& tab, const Fr* s, size_t n, uint8_t* out, int nthreads) {
  constexpr size_t PSZ = sizeof(F) == sizeof(Fq) ? 64 : 128;
  constexpr size_t CH = 2048;
  const size_t nchunks = (n + CH - 1) / CH;
  std::atomic<size_t> next(0);
  auto worker = [&] {
    std::vector<Jac<F>> j(CH);
    std::vector<Aff<F>> a(CH);
    for (;;) {
      const size_t c = next.fetch_add(1);
      if (c >= nchunks) return;
      const size_t lo = c * CH, cnt = std::min(CH, n - lo);
      for (size_t i = 0; i < cnt; i++) j[i] = tab.mul(s[lo + i]);
      batch_to_affine(j.data(), cnt, a.data());
      for (size_t i = 0; i < cnt; i++) encode(a[i], out + (lo + i) * PSZ);
    }
  };
  if (nthreads <= 1 || nchunks < 2) { worker(); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) th.emplace_back(worker);
  for (auto& t : th) t.join();
}

static Fq fq_small(u64 v) { Fq r = Fq::zero(); r.v[0] = v; return r.to_mont(); }
static Fq fq_limbs(u64 a, u64 b, u64 c, u64 d) { Fq r; r.v[0] = a; r.v[1] = b; r.v[2] = c; r.v[3] = d; return r.to_mont(); }

}  // namespace

extern "C" {

void* oracle_circuit_synth(u64 n_rows, u64 seed) {
  if (n_rows < 3 || n_rows > (1ull << 27)) return nullptr;
  SplitMix64 rng(seed);
  const uint32_t n_gates = (uint32_t)(n_rows - 2);
  OCircuit* c = new OCircuit();
  c->n_in = 2;
  c->n_gates = n_gates;
  std::vector<Fr>& aux = c->aux;
  aux.reserve(N_INIT_AUX + n_gates);
  for (int i = 0; i < N_INIT_AUX; i++) aux.push_back(rng.fr_mont());
  c->inputs = {Fr::one(), aux[0]};
  for (int m = 0; m < 3; m++) {
    c->rowptr[m].reserve(n_gates + 1);
    c->rowptr[m].push_back(0);
    const size_t cap = m == 2 ? n_gates : 3ull * n_gates;
    c->col[m].reserve(cap);
    c->coef[m].reserve(cap);
  }
  const Fr one = Fr::one();
  // row 0: inputize  [1*Aux0] * [1*Input0] = [1*Input1]
  c->col[0].push_back(2); c->coef[0].push_back(one);
  c->col[1].push_back(0); c->coef[1].push_back(one);
  c->col[2].push_back(1); c->coef[2].push_back(one);
  for (int m = 0; m < 3; m++) c->rowptr[m].push_back(1);
  for (uint32_t g = 1; g < n_gates; g++) {
    Fr ev[2];
    for (int side = 0; side < 2; side++) {
      Fr acc = Fr::zero();
      for (int t = 0; t < 3; t++) {
        const u64 u = rng.next() % (2 + aux.size());
        const Fr& val = u < 2 ? c->inputs[u] : aux[u - 2];
        c->col[side].push_back((uint32_t)u);
        if (rng.next() & 1) {
          c->coef[side].push_back(one);
          acc = acc + val;
        } else {
          const Fr cf = rng.fr_mont();
          c->coef[side].push_back(cf);
          acc = acc + cf * val;
        }
      }
      c->rowptr[side].push_back((uint32_t)c->col[side].size());
      ev[side] = acc;
    }
    aux.push_back(ev[0] * ev[1]);
    c->col[2].push_back(2 + (uint32_t)aux.size() - 1);
    c->coef[2].push_back(one);
    c->rowptr[2].push_back((uint32_t)c->col[2].size());
  }
  c->n_aux = (uint32_t)aux.size();
  return c;
}

// Circuit from caller arrays (tests: the same gates the Python oracle holds).  coef: Montgomery, one per term.
void* oracle_circuit_from_csr(uint32_t n_gates, uint32_t n_in, uint32_t n_aux, const uint32_t* const rowptr[3],
                              const uint32_t* const col[3], const u64* const coef[3]) {
  OCircuit* c = new OCircuit();
  c->n_in = n_in; c->n_aux = n_aux; c->n_gates = n_gates;
  for (int m = 0; m < 3; m++) {
    c->rowptr[m].assign(rowptr[m], rowptr[m] + n_gates + 1);
    const size_t nnz = rowptr[m][n_gates];
    c->col[m].assign(col[m], col[m] + nnz);
    c->coef[m].resize(nnz);
    if (nnz) memcpy(c->coef[m].data(), coef[m], nnz * 32);
  }
  return c;
}

void oracle_circuit_free(void* h) { delete reinterpret_cast<OCircuit*>(h); }

// shape[4] = n_in, n_aux, n_gates, nnz(total); arrays: pointers into the handle (valid until it is freed)
void oracle_circuit_view(void* h, u64 shape[4], const uint32_t* rowptr[3], const uint32_t* col[3], const u64* coef[3],
                         u64 nnz[3], const u64** inputs, const u64** aux) {
  OCircuit* c = reinterpret_cast<OCircuit*>(h);
  shape[0] = c->n_in; shape[1] = c->n_aux; shape[2] = c->n_gates;
  shape[3] = c->col[0].size() + c->col[1].size() + c->col[2].size();
  for (int m = 0; m < 3; m++) {
    rowptr[m] = c->rowptr[m].data();
    col[m] = c->col[m].data();
    coef[m] = reinterpret_cast<const u64*>(c->coef[m].data());
    nnz[m] = c->col[m].size();
  }
  *inputs = c->inputs.empty() ? nullptr : reinterpret_cast<const u64*>(c->inputs.data());
  *aux = c->aux.empty() ? nullptr : reinterpret_cast<const u64*>(c->aux.data());
}

// alpha beta gamma delta tau r s (Montgomery) from the stream seed ^ 0xB11D
void oracle_synth_trapdoor(u64 seed, u64 out[7][4]) {
  SplitMix64 rng(seed ^ 0xB11D);
  for (int i = 0; i < 7; i++) {
    Fr h = rng.fr_mont();
    memcpy(out[i], h.v, 32);
  }
}

// trapdoor: alpha beta gamma delta tau, Montgomery.  *params_out is malloc'ed (oracle_free).  stage_s optional
// [3]: scalar side, point side, total seconds.  Returns 0, or <0 on bad input.
int oracle_setup(void* h, const u64 trapdoor[5][4], int nthreads, uint8_t** params_out, size_t* len_out,
                 double* stage_s) {
  auto t0 = std::chrono::steady_clock::now();
  OCircuit* c = reinterpret_cast<OCircuit*>(h);
  if (!c || !trapdoor || !params_out || !len_out) return -1;
  if (nthreads < 1) nthreads = 1;
  const uint32_t n_in = c->n_in, n_aux = c->n_aux, ng = c->n_gates;
  const size_t n_rows = (size_t)ng + n_in;
  size_t m = 1; int exp = 0;
  while (m < n_rows) { m *= 2; exp++; if (exp >= 28) return -4; }
  if (exp == 0) { m = 2; exp = 1; }
  Fr alpha, beta, gamma, delta, tau;
  memcpy(alpha.v, trapdoor[0], 32); memcpy(beta.v, trapdoor[1], 32); memcpy(gamma.v, trapdoor[2], 32);
  memcpy(delta.v, trapdoor[3], 32); memcpy(tau.v, trapdoor[4], 32);
  if (gamma.is_zero() || delta.is_zero()) return -2;
  // powers of tau (chunked: every thread starts from tau^lo), h scalars
  std::vector<Fr> pw(m);
  parallel_for(nthreads, m, [&](size_t lo, size_t hi) {
    u64 e = lo;
    Fr u = tau.pow(&e, 1);
    for (size_t i = lo; i < hi; i++) { pw[i] = u; u = u * tau; }
  });
  const Fr z_tau = pw[m - 1] * tau - Fr::one();
  const Fr dinv = delta.inv(), ginv = gamma.inv();
  const Fr hcoef = z_tau * dinv;
  std::vector<Fr> h_s(m - 1);
  parallel_for(nthreads, m - 1, [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; i++) h_s[i] = pw[i] * hcoef; });
  // Lagrange values at tau: ifft of the powers
  {
    Fr omega; memcpy(omega.v, ROOT_OF_UNITY, 32);
    for (int i = exp; i < 28; i++) omega = omega.sqr();
    Fr mf = Fr::zero(); mf.v[0] = m; mf = mf.to_mont();
    fft(pw, omega.inv(), exp, nthreads);
    scale(pw, mf.inv(), nthreads);
  }
  const std::vector<Fr>& lag = pw;
  // column accumulation: acc[mi][k] = sum_rows M[row, k] L_row(tau); the three matrices on their own threads
  const uint32_t nv = n_in + n_aux;
  std::vector<Fr> acc[3];
  {
    std::vector<std::function<void()>> tasks;
    for (int mi = 0; mi < 3; mi++)
      tasks.push_back([&, mi] {
        acc[mi].assign(nv, Fr::zero());
        const Fr one = Fr::one();
        for (uint32_t row = 0; row < ng; row++) {
          const Fr& lj = lag[row];
          for (uint32_t p = c->rowptr[mi][row]; p < c->rowptr[mi][row + 1]; p++) {
            Fr& dst = acc[mi][c->col[mi][p]];
            const Fr& cf = c->coef[mi][p];
            dst = dst + (cf == one ? lj : cf * lj);
          }
        }
      });
    run_tasks(std::min(nthreads, 3), tasks);
  }
  for (uint32_t i = 0; i < n_in; i++) acc[0][i] = acc[0][i] + lag[ng + i];  // input_i * 0 = 0 rows
  std::vector<Fr> ic_s(n_in), l_s(n_aux), a_s, b_s;
  parallel_for(nthreads, nv, [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) {
      const Fr t = beta * acc[0][i] + alpha * acc[1][i] + acc[2][i];
      if (i < n_in) ic_s[i] = t * ginv;
      else l_s[i - n_in] = t * dinv;
    }
  });
  for (uint32_t i = 0; i < nv; i++) {  // points at infinity are filtered out of the a / b queries
    if (!acc[0][i].is_zero()) a_s.push_back(acc[0][i]);
    if (!acc[1][i].is_zero()) b_s.push_back(acc[1][i]);
  }
  for (int mi = 0; mi < 3; mi++) std::vector<Fr>().swap(acc[mi]);
  std::vector<Fr>().swap(pw);
  auto t1 = std::chrono::steady_clock::now();
  // ---- point side
  const size_t n_h = m - 1, n_a = a_s.size(), n_b = b_s.size();
  const size_t total = 64 + 64 + 128 + 128 + 64 + 128 + 4 + (size_t)n_in * 64 + 4 + n_h * 64 + 4 + (size_t)n_aux * 64 +
                       4 + n_a * 64 + 4 + n_b * 64 + 4 + n_b * 128;
  uint8_t* out = (uint8_t*)malloc(total);
  if (!out) return -3;
  FixedBaseTable<Fq> t1g;
  FixedBaseTable<Fq2> t2g;
  t1g.build({fq_small(1), fq_small(2), false}, nthreads);
  t2g.build({{fq_limbs(0x46debd5cd992f6edull, 0x674322d4f75edaddull, 0x426a00665e5c4479ull, 0x1800deef121f1e76ull),
              fq_limbs(0x97e485b7aef312c2ull, 0xf1aa493335a9e712ull, 0x7260bfb731fb5d25ull, 0x198e9393920d483aull)},
             {fq_limbs(0x4ce6cc0166fa7daaull, 0xe3d1e7690c43d37bull, 0x4aab71808dcb408full, 0x12c85ea5db8c6debull),
              fq_limbs(0x55acdadcd122975bull, 0xbc4b313370b38ef3ull, 0xec9e99ad690c3395ull, 0x090689d0585ff075ull)},
             false},
            nthreads);
  size_t pos = 0;
  auto len_be = [&](u64 n) {
    out[pos] = (uint8_t)(n >> 24); out[pos + 1] = (uint8_t)(n >> 16); out[pos + 2] = (uint8_t)(n >> 8); out[pos + 3] = (uint8_t)n;
    pos += 4;
  };
  auto emit1 = [&](const Fr* s, size_t n) { emit_points(t1g, s, n, out + pos, nthreads); pos += n * 64; };
  auto emit2 = [&](const Fr* s, size_t n) { emit_points(t2g, s, n, out + pos, nthreads); pos += n * 128; };
  emit1(&alpha, 1); emit1(&beta, 1); emit2(&beta, 1); emit2(&gamma, 1); emit1(&delta, 1); emit2(&delta, 1);
  len_be(n_in); emit1(ic_s.data(), n_in);
  len_be(n_h); emit1(h_s.data(), n_h);
  len_be(n_aux); emit1(l_s.data(), n_aux);
  len_be(n_a); emit1(a_s.data(), n_a);
  len_be(n_b); emit1(b_s.data(), n_b);
  len_be(n_b); emit2(b_s.data(), n_b);
  *params_out = out;
  *len_out = total;
  auto t2 = std::chrono::steady_clock::now();
  if (stage_s) {
    auto sec = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
    stage_s[0] = sec(t0, t1); stage_s[1] = sec(t1, t2); stage_s[2] = sec(t0, t2);
  }
  return 0;
}

void oracle_free(void* p) { free(p); }

}  // extern "C"
