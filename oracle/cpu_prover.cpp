// CPU restatement of the reference's Groth16 prover (bellman_ce 0.3.5 shape) -- C++17, threads.
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; the product
// (fawkes-crypto_b200/) never does.  It shares no code with the product: fields are 4 x 64-bit
// limbs with unsigned __int128, points are Jacobian, transforms are plain radix-2 with
// bit reversal, multiexp uses bellman's unsigned windows -- deliberately different from the
// CUDA path so that byte-equal proofs are meaningful.
//
// PARITY STATUS: "parity unpinned" at proof level -- the reference prover body lives in
// crates absent from /root/reference (fawkes-crypto-bellman_ce 0.3.5, pairing_ce 0.18.1;
// Cargo.lock:413-436) and the reference has no golden proofs.  This file restates their
// published algorithm and is itself checked against oracle/groth16.py
// (tests/test_oracle.py), which is anchored by an independent pairing check.
//
// What each part follows:
//   field mul/add/sub        ff-uint_derive/src/lib.rs:434-490,578-623,836-862 (same results)
//   moduli                   fawkes-crypto/src/engines/bn256/mod.rs:13,23
//   rows / variables         fawkes-crypto/src/backend/bellman_groth16/mod.rs:61-102
//   eval, density, domain, multiexp (c = ln n, one task per window), assembly
//                            bellman_ce prover.rs / domain.rs / multiexp.rs [App. C.1-C.5]
//   point byte formats       SURVEY.md App. B
#include "cpu_common.h"

// -------------------------------------------------------------- multiexp ---
// bellman: c = n < 32 ? 3 : ceil(ln n); unsigned windows; scalar 0 skipped, scalar 1 added
// directly in the first window; one task per window; windows joined by c doublings.
template <class F>
struct Multiexp {
  const Aff<F>* bases;
  const u64* exps;  // canonical, 4 limbs each
  size_t n;
  int c, nwin;
  std::vector<Jac<F>> win;
  void plan(const Aff<F>* b, const u64* e, size_t n_) {
    bases = b; exps = e; n = n_;
    c = n < 32 ? 3 : (int)std::ceil(std::log((double)n));
    nwin = (254 + c - 1) / c;  // skip = 0, c, 2c, ... < 254 (Fr::NUM_BITS)
    win.assign(nwin, Jac<F>::zero());
  }
  void run_window(int w) {
    const int skip = w * c;
    Jac<F> acc = Jac<F>::zero();
    std::vector<Jac<F>> buckets((size_t(1) << c) - 1, Jac<F>::zero());
    const u64 mask = (u64(1) << c) - 1;
    for (size_t i = 0; i < n; i++) {
      const u64* e = exps + 4 * i;
      if ((e[0] | e[1] | e[2] | e[3]) == 0) continue;
      if (e[0] == 1 && (e[1] | e[2] | e[3]) == 0) {
        if (w == 0) acc.add_mixed(bases[i]);
        continue;
      }
      const int limb = skip >> 6, sh = skip & 63;
      u64 d = e[limb] >> sh;
      if (sh && limb + 1 < 4) d |= e[limb + 1] << (64 - sh);
      d &= mask;
      if (d) buckets[d - 1].add_mixed(bases[i]);
    }
    Jac<F> run = Jac<F>::zero();
    for (size_t b = buckets.size(); b-- > 0;) {
      run.add(buckets[b]);
      acc.add(run);
    }
    win[w] = acc;
  }
  Jac<F> finish() {
    Jac<F> r = Jac<F>::zero();
    for (int w = nwin - 1; w >= 0; w--) {
      for (int i = 0; i < c; i++) r.dbl();
      r.add(win[w]);
    }
    return r;
  }
};

extern "C" {

int oracle_hw_threads() { return (int)std::thread::hardware_concurrency(); }

// Full prover.  params = bellman Parameters bytes.  Matrices in CSR over w = [inputs|aux]:
// rowptr[m] (n_gates+1), col[m], coef[m] (Montgomery Fr, 4 x u64 per term).  inputs/aux/r/s:
// Montgomery Fr.  proof_raw: 256 B (a.x a.y | b.x.c0 b.x.c1 b.y.c0 b.y.c1 | c.x c.y raw LE
// Montgomery, zeros = infinity).  h_out optional [m-1][4].  stage_s optional [4]:
// eval, fft, multiexp, total seconds.  Returns 0, or <0 on a malformed input.
int oracle_groth16_prove(const uint8_t* params, size_t plen, uint32_t n_gates, uint32_t n_in, uint32_t n_aux,
                         const uint32_t* const rowptr[3], const uint32_t* const col[3], const u64* const coef[3],
                         const u64* inputs, const u64* aux, const u64* r_, const u64* s_, int nthreads,
                         uint8_t* proof_raw, u64* h_out, double* stage_s) {
  auto t_start = std::chrono::steady_clock::now();
  if (nthreads < 1) nthreads = 1;
  // ---- parse parameters
  if (plen < 580) return -1;
  size_t pos = 0;
  Aff<Fq> alpha_g1 = g1_from_be(params); pos += 64;
  Aff<Fq> beta_g1 = g1_from_be(params + pos); pos += 64;
  Aff<Fq2> beta_g2 = g2_from_be(params + pos); pos += 128;
  pos += 128;  // gamma_g2
  Aff<Fq> delta_g1 = g1_from_be(params + pos); pos += 64;
  Aff<Fq2> delta_g2 = g2_from_be(params + pos); pos += 128;
  uint32_t n_ic = be32(params + pos); pos += 4 + (size_t)n_ic * 64;
  if (n_ic != n_in) return -2;
  std::vector<Aff<Fq>> hq, lq, aq, b1q;
  std::vector<Aff<Fq2>> b2q;
  auto read_g1 = [&](std::vector<Aff<Fq>>& v) -> bool {
    if (pos + 4 > plen) return false;
    uint32_t n = be32(params + pos); pos += 4;
    if (pos + (size_t)n * 64 > plen) return false;
    v.resize(n);
    const uint8_t* base = params + pos;
    parallel_for(nthreads, n, [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; i++) v[i] = g1_from_be(base + i * 64); });
    pos += (size_t)n * 64;
    return true;
  };
  if (!read_g1(hq) || !read_g1(lq) || !read_g1(aq) || !read_g1(b1q)) return -1;
  {
    if (pos + 4 > plen) return -1;
    uint32_t n = be32(params + pos); pos += 4;
    if (pos + (size_t)n * 128 > plen) return -1;
    b2q.resize(n);
    const uint8_t* base = params + pos;
    parallel_for(nthreads, n, [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; i++) b2q[i] = g2_from_be(base + i * 128); });
    pos += (size_t)n * 128;
  }
  if (lq.size() != n_aux) return -2;
  auto t_parse = std::chrono::steady_clock::now();

  // ---- eval (ProvingAssignment::enforce) + density
  const size_t n_rows = (size_t)n_gates + n_in;
  size_t m = 1; int exp = 0;
  while (m < n_rows) { m *= 2; exp++; if (exp >= 28) return -4; }
  if (exp == 0) { m = 2; exp = 1; }
  const Fr* W_in = reinterpret_cast<const Fr*>(inputs);
  const Fr* W_aux = reinterpret_cast<const Fr*>(aux);
  auto wv = [&](uint32_t c) -> const Fr& { return c < n_in ? W_in[c] : W_aux[c - n_in]; };
  std::vector<Fr> ev[3];
  std::vector<uint8_t> a_dens(n_in + n_aux, 0), b_dens(n_in + n_aux, 0);
  for (int mi = 0; mi < 3; mi++) {
    ev[mi].assign(m, Fr::zero());
    parallel_for(nthreads, n_gates, [&](size_t lo, size_t hi) {
      const Fr one = Fr::one();
      for (size_t row = lo; row < hi; row++) {
        Fr acc = Fr::zero();
        for (uint32_t p = rowptr[mi][row]; p < rowptr[mi][row + 1]; p++) {
          const Fr& cf = *reinterpret_cast<const Fr*>(coef[mi] + 4 * (size_t)p);
          const uint32_t c = col[mi][p];
          if (mi == 0) a_dens[c] = 1;
          if (mi == 1) b_dens[c] = 1;
          if (cf == one) acc = acc + wv(c);
          else acc = acc + wv(c) * cf;
        }
        ev[mi][row] = acc;
      }
    });
  }
  for (uint32_t i = 0; i < n_in; i++) ev[0][n_gates + i] = W_in[i];  // input_i * 0 = 0
  auto t_eval = std::chrono::steady_clock::now();

  // ---- H
  Fr omega; memcpy(omega.v, ROOT_OF_UNITY, 32);
  for (int i = exp; i < 28; i++) omega = omega.sqr();
  Fr g; memcpy(g.v, GEN7, 32);
  Fr omega_inv = omega.inv(), g_inv = g.inv();
  Fr mf = Fr::zero(); mf.v[0] = m; mf = mf.to_mont();
  Fr minv = mf.inv();
  for (int mi = 0; mi < 3; mi++) {
    fft(ev[mi], omega_inv, exp, nthreads);       // ifft
    scale(ev[mi], minv, nthreads);
    distribute_powers(ev[mi], g, nthreads);      // coset_fft
    fft(ev[mi], omega, exp, nthreads);
  }
  Fr zinv = (fr_pow_u64(g, m) - Fr::one()).inv();
  parallel_for(nthreads, m, [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) ev[0][i] = (ev[0][i] * ev[1][i] - ev[2][i]) * zinv;
  });
  fft(ev[0], omega_inv, exp, nthreads);          // icoset_fft
  scale(ev[0], minv, nthreads);
  distribute_powers(ev[0], g_inv, nthreads);
  std::vector<Fr>& h = ev[0];
  h.resize(m - 1);
  if (h_out) memcpy(h_out, h.data(), (m - 1) * 32);
  auto t_fft = std::chrono::steady_clock::now();

  // ---- scalars in canonical form, density selection
  std::vector<u64> h_rep((m - 1) * 4), in_rep((size_t)n_in * 4), aux_rep((size_t)n_aux * 4);
  parallel_for(nthreads, m - 1, [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; i++) { Fr t = h[i].from_mont(); memcpy(&h_rep[4 * i], t.v, 32); } });
  parallel_for(nthreads, n_aux, [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; i++) { Fr t = W_aux[i].from_mont(); memcpy(&aux_rep[4 * i], t.v, 32); } });
  for (uint32_t i = 0; i < n_in; i++) { Fr t = W_in[i].from_mont(); memcpy(&in_rep[4 * i], t.v, 32); }
  std::vector<u64> a_aux_rep, b_in_rep, b_aux_rep;
  for (uint32_t i = 0; i < n_aux; i++) {
    if (a_dens[n_in + i]) a_aux_rep.insert(a_aux_rep.end(), &aux_rep[4 * i], &aux_rep[4 * i] + 4);
    if (b_dens[n_in + i]) b_aux_rep.insert(b_aux_rep.end(), &aux_rep[4 * i], &aux_rep[4 * i] + 4);
  }
  for (uint32_t i = 0; i < n_in; i++)
    if (b_dens[i]) b_in_rep.insert(b_in_rep.end(), &in_rep[4 * i], &in_rep[4 * i] + 4);
  const size_t nb_in = b_in_rep.size() / 4, nb_aux = b_aux_rep.size() / 4, na_aux = a_aux_rep.size() / 4;
  if (hq.size() < m - 1) return -3;
  if (aq.size() != n_in + na_aux || b1q.size() != nb_in + nb_aux || b2q.size() != b1q.size()) return -6;

  // ---- the eight multiexps, one task per window (bellman Worker shape)
  Multiexp<Fq> mh, ml, ma_in, ma_aux, mb1_in, mb1_aux;
  Multiexp<Fq2> mb2_in, mb2_aux;
  mh.plan(hq.data(), h_rep.data(), m - 1);
  ml.plan(lq.data(), aux_rep.data(), n_aux);
  ma_in.plan(aq.data(), in_rep.data(), n_in);
  ma_aux.plan(aq.data() + n_in, a_aux_rep.data(), na_aux);
  mb1_in.plan(b1q.data(), b_in_rep.data(), nb_in);
  mb1_aux.plan(b1q.data() + nb_in, b_aux_rep.data(), nb_aux);
  mb2_in.plan(b2q.data(), b_in_rep.data(), nb_in);
  mb2_aux.plan(b2q.data() + nb_in, b_aux_rep.data(), nb_aux);
  std::vector<std::function<void()>> tasks;
  auto push1 = [&](Multiexp<Fq>& x) { for (int w = 0; w < x.nwin; w++) tasks.push_back([&x, w] { x.run_window(w); }); };
  auto push2 = [&](Multiexp<Fq2>& x) { for (int w = 0; w < x.nwin; w++) tasks.push_back([&x, w] { x.run_window(w); }); };
  push2(mb2_aux); push2(mb2_in);  // longest first
  push1(mh); push1(ml); push1(ma_aux); push1(mb1_aux); push1(ma_in); push1(mb1_in);
  run_tasks(nthreads, tasks);
  auto t_msm = std::chrono::steady_clock::now();

  // ---- assembly (App. C.5)
  if (delta_g1.inf || delta_g2.inf) return -5;
  Fr r, s; memcpy(r.v, r_, 32); memcpy(s.v, s_, 32);
  Fr rc = r.from_mont(), sc = s.from_mont(), rsc = (r * s).from_mont();
  auto J1 = [](const Aff<Fq>& p) { Jac<Fq> j = Jac<Fq>::zero(); j.add_mixed(p); return j; };
  auto J2 = [](const Aff<Fq2>& p) { Jac<Fq2> j = Jac<Fq2>::zero(); j.add_mixed(p); return j; };
  Jac<Fq> g_a = J1(delta_g1); g_a.mul_assign(rc.v); g_a.add_mixed(alpha_g1);
  Jac<Fq2> g_b = J2(delta_g2); g_b.mul_assign(sc.v); g_b.add_mixed(beta_g2);
  Jac<Fq> g_c = J1(delta_g1); g_c.mul_assign(rsc.v);
  { Jac<Fq> t = J1(alpha_g1); t.mul_assign(sc.v); g_c.add(t); }
  { Jac<Fq> t = J1(beta_g1); t.mul_assign(rc.v); g_c.add(t); }
  Jac<Fq> a_ans = ma_in.finish(); a_ans.add(ma_aux.finish());
  g_a.add(a_ans);
  a_ans.mul_assign(sc.v);
  g_c.add(a_ans);
  Jac<Fq> b1_ans = mb1_in.finish(); b1_ans.add(mb1_aux.finish());
  Jac<Fq2> b2_ans = mb2_in.finish(); b2_ans.add(mb2_aux.finish());
  g_b.add(b2_ans);
  b1_ans.mul_assign(rc.v);
  g_c.add(b1_ans);
  g_c.add(mh.finish());
  g_c.add(ml.finish());
  Aff<Fq> pa = g_a.to_affine(), pc = g_c.to_affine();
  Aff<Fq2> pb = g_b.to_affine();
  memset(proof_raw, 0, 256);
  if (!pa.inf) { memcpy(proof_raw, pa.x.v, 32); memcpy(proof_raw + 32, pa.y.v, 32); }
  if (!pb.inf) {
    memcpy(proof_raw + 64, pb.x.c0.v, 32); memcpy(proof_raw + 96, pb.x.c1.v, 32);
    memcpy(proof_raw + 128, pb.y.c0.v, 32); memcpy(proof_raw + 160, pb.y.c1.v, 32);
  }
  if (!pc.inf) { memcpy(proof_raw + 192, pc.x.v, 32); memcpy(proof_raw + 224, pc.y.v, 32); }
  auto t_end = std::chrono::steady_clock::now();
  if (stage_s) {
    auto sec = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
    stage_s[0] = sec(t_parse, t_eval);
    stage_s[1] = sec(t_eval, t_fft);
    stage_s[2] = sec(t_fft, t_msm);
    stage_s[3] = sec(t_parse, t_end);  // prove time excludes parameter parsing (load-time in the reference)
  }
  (void)t_start;
  return 0;
}

// single multiexp, for the G1 MSM Mpoints/s line: bases raw affine Montgomery (64 B), scalars
// Montgomery Fr.  result raw affine.  Returns seconds.
double oracle_msm_g1(const uint8_t* bases_raw, const u64* scalars, size_t n, int nthreads, uint8_t* result_raw) {
  std::vector<Aff<Fq>> b(n);
  std::vector<u64> e(n * 4);
  for (size_t i = 0; i < n; i++) {
    memcpy(b[i].x.v, bases_raw + 64 * i, 32);
    memcpy(b[i].y.v, bases_raw + 64 * i + 32, 32);
    b[i].inf = b[i].x.is_zero() && b[i].y.is_zero();
    Fr t; memcpy(t.v, scalars + 4 * i, 32); t = t.from_mont(); memcpy(&e[4 * i], t.v, 32);
  }
  auto t0 = std::chrono::steady_clock::now();
  Multiexp<Fq> mx;
  mx.plan(b.data(), e.data(), n);
  std::vector<std::function<void()>> tasks;
  for (int w = 0; w < mx.nwin; w++) tasks.push_back([&mx, w] { mx.run_window(w); });
  run_tasks(nthreads, tasks);
  Aff<Fq> res = mx.finish().to_affine();
  auto t1 = std::chrono::steady_clock::now();
  memset(result_raw, 0, 64);
  if (!res.inf) { memcpy(result_raw, res.x.v, 32); memcpy(result_raw + 32, res.y.v, 32); }
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
