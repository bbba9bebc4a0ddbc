// Groth16 prover on one B200: R1CS evaluation -> H pipeline -> five MSMs -> assembly.
//
// Replaces bellman::groth16::create_random_proof as called from
// fawkes-crypto/src/backend/bellman_groth16/prover.rs:78-80, including the work
// BellmanCS::synthesize streams into it (backend/bellman_groth16/mod.rs:61-102):
//   ProvingAssignment::enforce/eval  -> k_spmv (a_i = <A_i,w>, b_i, c_i) + input rows
//   EvaluationDomain pipeline        -> NttDomain (ntt.cu)
//   8 multiexps                      -> 5 MSMs (msm.cu); the *_inputs / *_aux pairs are
//                                       merged because the query arrays are contiguous
//   final assembly                   -> host, SURVEY.md App. C.5
#include <chrono>
#include <cstring>

#include "internal.h"

namespace fb {

// --------------------------------------------------------------- R1CS eval ---
// One thread per row.  cidx: 0 -> +w, 1 -> -w, k -> coef[k-2]*w
// (bellman's eval skips the multiply when coeff == 1; -1 is our addition.)
__global__ void k_spmv(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ col,
                       const uint32_t* __restrict__ cidx, const Fr* __restrict__ coef,
                       const Fr* __restrict__ w, Fr* __restrict__ out, uint32_t n_rows) {
  for (uint32_t row = blockIdx.x * blockDim.x + threadIdx.x; row < n_rows;
       row += gridDim.x * blockDim.x) {
    Fr acc = Fr::zero();
    const uint32_t e = rowptr[row + 1];
    for (uint32_t p = rowptr[row]; p < e; p++) {
      Fr v = w[col[p]];
      const uint32_t ci = cidx[p];
      if (ci == 0) acc = add(acc, v);
      else if (ci == 1) acc = sub(acc, v);
      else acc = add(acc, mul(v, coef[ci - 2]));
    }
    out[row] = acc;
  }
}

int eval_r1cs(const DevCsr& csr, const Fr* w, uint32_t n_in, Fr* a, Fr* b, Fr* c, uint64_t m,
              cudaStream_t st) {
  const uint32_t ng = csr.n_gates;
  Fr* outs[3] = {a, b, c};
  for (int i = 0; i < 3; i++) {
    // rows >= n_gates: zero (b, c) -- a gets the input rows below
    FB_CUDA(cudaMemsetAsync(outs[i] + ng, 0, (m - ng) * sizeof(Fr), st));
    if (ng) {
      unsigned blocks = (unsigned)std::min<uint64_t>((ng + 127) / 128, 148 * 32);
      k_spmv<<<blocks, 128, 0, st>>>(csr.rowptr[i], csr.col[i], csr.cidx[i], csr.coef, w, outs[i], ng);
      count_launch();
    }
  }
  // bellman appends `input_i * 0 = 0` for every input: a = w_i, b = c = 0
  FB_CUDA(cudaMemcpyAsync(a + ng, w, (size_t)n_in * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
  return FB_OK;
}

// Distributed variant: this rank holds the rows i = rank (mod 2^g) (local row j <-> global row
// j * 2^g + rank), the layout the distributed H pipeline starts from.
int eval_r1cs_cyclic(const DevCsr& csr, const Fr* w, uint32_t n_in, uint32_t n_gates_global, int g, int rank,
                     Fr* a, Fr* b, Fr* c, uint64_t ml, cudaStream_t st) {
  const uint32_t ng = csr.n_gates;  // local rows
  Fr* outs[3] = {a, b, c};
  for (int i = 0; i < 3; i++) {
    FB_CUDA(cudaMemsetAsync(outs[i] + ng, 0, (ml - ng) * sizeof(Fr), st));
    if (ng) {
      unsigned blocks = (unsigned)std::min<uint64_t>((ng + 127) / 128, 148 * 32);
      k_spmv<<<blocks, 128, 0, st>>>(csr.rowptr[i], csr.col[i], csr.cidx[i], csr.coef, w, outs[i], ng);
      count_launch();
    }
  }
  // bellman's `input_i * 0 = 0` rows: global row n_gates + i
  for (uint32_t i = 0; i < n_in; i++) {
    const uint64_t row = (uint64_t)n_gates_global + i;
    if ((int)(row & ((1u << g) - 1)) == rank)
      FB_CUDA(cudaMemcpyAsync(a + (row >> g), w + i, sizeof(Fr), cudaMemcpyDeviceToDevice, st));
  }
  return FB_OK;
}

}  // namespace fb
