// Synthetic random R1CS + witness of SURVEY.md section 8(d) (benchmark / test harness).
//
// Shape mimics the reference's circuit builder: row 0 is BuildCS::inputize
// (fawkes-crypto/src/circuit/r1cs/cs.rs:309-318), every other row is one CNum
// multiplication (fawkes-crypto/src/circuit/r1cs/num.rs:253-272): A_i, B_i are 3-term
// combinations of earlier variables (coefficient 1 w.p. 1/2 else uniform Fr), C_i is the
// freshly allocated product.  PRNG = SplitMix64; Fr sample = 4 words, top limb masked to
// 62 bits, reject >= r.  oracle/synth.py is the Python restatement; tests check both
// produce identical circuits and witnesses.
#include "../../include/fawkes_b200.h"

#include "host_fr.h"
#include <memory>

#include "internal.h"

namespace fb {

struct SplitMix64 {
  uint64_t s;
  explicit SplitMix64(uint64_t seed) : s(seed) {}
  uint64_t next() {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  hfr::H fr_canonical() {
    for (;;) {
      hfr::H h;
      for (int i = 0; i < 4; i++) h.v[i] = next();
      h.v[3] &= (1ull << 62) - 1;
      if (!hfr::geq_mod(h.v)) return h;
    }
  }
  hfr::H fr_mont() { return hfr::to_mont(fr_canonical()); }
};

static const int N_INIT_AUX = 16;

static inline Fr to_dev(const hfr::H& h) {
  Fr r;
  memcpy(r.v, h.v, 32);
  return r;
}

}  // namespace fb

using namespace fb;

extern "C" {

int fb_circuit_synth(uint64_t n_rows, uint64_t seed, fb_circuit** out) try {
  if (!out || n_rows < 3 || n_rows > (1ull << 27)) { set_error("fb_circuit_synth: bad size"); return FB_ERR_ARG; }
  SplitMix64 rng(seed);
  const uint32_t n_gates = (uint32_t)(n_rows - 2);
  std::unique_ptr<Circuit> c(new Circuit());
  c->n_in = 2;
  std::vector<hfr::H> aux;
  aux.reserve(N_INIT_AUX + n_gates);
  for (int i = 0; i < N_INIT_AUX; i++) aux.push_back(rng.fr_mont());
  hfr::H inputs[2] = {hfr::one(), aux[0]};
  HostCsr& csr = c->csr;
  for (int m = 0; m < 3; m++) {
    csr.rowptr[m].reserve(n_gates + 1);
    csr.rowptr[m].push_back(0);
    csr.col[m].reserve(m == 2 ? n_gates : 3ull * n_gates);
    csr.cidx[m].reserve(m == 2 ? n_gates : 3ull * n_gates);
  }
  csr.coef.reserve(3ull * n_gates + 16);
  // row 0: inputize  [1*Aux0] * [1*Input0] = [1*Input1]
  csr.col[0].push_back(2 + 0); csr.cidx[0].push_back(0);
  csr.col[1].push_back(0);     csr.cidx[1].push_back(0);
  csr.col[2].push_back(1);     csr.cidx[2].push_back(0);
  for (int m = 0; m < 3; m++) csr.rowptr[m].push_back(1);
  for (uint32_t g = 1; g < n_gates; g++) {
    hfr::H ev[2];
    for (int side = 0; side < 2; side++) {
      hfr::H acc = hfr::zero();
      for (int t = 0; t < 3; t++) {
        const uint64_t u = rng.next() % (2 + aux.size());
        const hfr::H& val = u < 2 ? inputs[u] : aux[u - 2];
        csr.col[side].push_back((uint32_t)u);
        if (rng.next() & 1) {
          csr.cidx[side].push_back(0);
          acc = hfr::add(acc, val);
        } else {
          hfr::H cf = rng.fr_mont();
          csr.cidx[side].push_back((uint32_t)csr.coef.size() + 2);
          csr.coef.push_back(to_dev(cf));
          acc = hfr::add(acc, hfr::mul(cf, val));
        }
      }
      csr.rowptr[side].push_back((uint32_t)csr.col[side].size());
      ev[side] = acc;
    }
    aux.push_back(hfr::mul(ev[0], ev[1]));
    csr.col[2].push_back(2 + (uint32_t)aux.size() - 1);
    csr.cidx[2].push_back(0);
    csr.rowptr[2].push_back((uint32_t)csr.col[2].size());
  }
  csr.n_gates = n_gates;
  c->n_aux = (uint32_t)aux.size();
  c->inputs = {to_dev(inputs[0]), to_dev(inputs[1])};
  c->aux.resize(aux.size());
  memcpy(c->aux.data(), aux.data(), aux.size() * 32);
  *out = reinterpret_cast<fb_circuit*>(c.release());
  return FB_OK;
} FB_ABI_CATCH_INT

int fb_circuit_witness(const fb_circuit* c_, const uint64_t** inputs, const uint64_t** aux) try {
  const Circuit* c = reinterpret_cast<const Circuit*>(c_);
  if (!c || c->inputs.empty()) { set_error("circuit carries no witness"); return FB_ERR_ARG; }
  if (inputs) *inputs = reinterpret_cast<const uint64_t*>(c->inputs.data());
  if (aux) *aux = reinterpret_cast<const uint64_t*>(c->aux.data());
  return FB_OK;
} FB_ABI_CATCH_INT

int fb_synth_trapdoor(uint64_t seed, uint64_t out[7][4]) try {
  if (!out) return FB_ERR_ARG;
  SplitMix64 rng(seed ^ 0xB11D);
  for (int i = 0; i < 7; i++) {
    hfr::H h = rng.fr_mont();
    memcpy(out[i], h.v, 32);
  }
  return FB_OK;
} FB_ABI_CATCH_INT

}  // extern "C"
