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
// LANES consecutive lanes per row (1 or 8).  cidx: 0 -> +w, 1 -> -w, k -> coef[k-2]*w
// (bellman's eval skips the multiply when coeff == 1; -1 is our addition.)
// LANES = 1 suits the 1-3 term rows of multiplication gates; gadget circuits (Poseidon's MDS rows carry up to 55
// terms) take 8 lanes per row: the terms of a row are strided over the lanes and the partial sums folded with
// three shuffle steps, so a long row no longer serialises 55 dependent multiply-adds in one thread.
template <int LANES>
__global__ void k_spmv(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ col,
                       const uint32_t* __restrict__ cidx, const Fr* __restrict__ coef,
                       const Fr* __restrict__ w, Fr* __restrict__ out, uint32_t n_rows, uint64_t wstride,
                       uint64_t ostride) {
  // batched proves: blockIdx.y selects one of gridDim.y witnesses / output arrays (strides in elements)
  w += (uint64_t)blockIdx.y * wstride;
  out += (uint64_t)blockIdx.y * ostride;
  const uint32_t lane_in_row = threadIdx.x % LANES;
  // whole lane groups stay together: the loop bound is rounded up so that every lane of a group takes part in the shuffles
  const uint32_t groups = gridDim.x * blockDim.x / LANES;
  for (uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) / LANES; row < ((n_rows + groups - 1) / groups) * groups;
       row += groups) {
    Fr acc = Fr::zero();
    if (row < n_rows) {
      const uint32_t e = rowptr[row + 1];
      for (uint32_t p = rowptr[row] + lane_in_row; p < e; p += LANES) {
        Fr v = w[col[p]];
        const uint32_t ci = cidx[p];
        if (ci == 0) acc = add(acc, v);
        else if (ci == 1) acc = sub(acc, v);
        else acc = add(acc, mul(v, coef[ci - 2]));
      }
    }
    if (LANES > 1) {
#pragma unroll
      for (int d = LANES / 2; d > 0; d >>= 1) {
        Fr o;
#pragma unroll
        for (int i = 0; i < 8; i++) o.v[i] = __shfl_down_sync(0xffffffffu, acc.v[i], d, LANES);
        acc = add(acc, o);
      }
    }
    if (row < n_rows && lane_in_row == 0) out[row] = acc;
  }
}

static void launch_spmv(const DevCsr& csr, int i, const Fr* w, Fr* out, uint64_t wstride, uint64_t ostride, unsigned count,
                        cudaStream_t st) {
  const uint32_t ng = csr.n_gates;
  if (csr.nnz[i] >= 4ull * ng) {
    const unsigned blocks = (unsigned)std::min<uint64_t>(((uint64_t)ng * 8 + 127) / 128, 148 * 32);
    k_spmv<8><<<dim3(blocks, count), 128, 0, st>>>(csr.rowptr[i], csr.col[i], csr.cidx[i], csr.coef, w, out, ng, wstride, ostride);
  } else {
    const unsigned blocks = (unsigned)std::min<uint64_t>((ng + 127) / 128, 148 * 32);
    k_spmv<1><<<dim3(blocks, count), 128, 0, st>>>(csr.rowptr[i], csr.col[i], csr.cidx[i], csr.coef, w, out, ng, wstride, ostride);
  }
  count_launch();
}

int eval_r1cs(const DevCsr& csr, const Fr* w, uint32_t n_in, Fr* a, Fr* b, Fr* c, uint64_t m,
              cudaStream_t st) {
  const uint32_t ng = csr.n_gates;
  Fr* outs[3] = {a, b, c};
  for (int i = 0; i < 3; i++) {
    // rows >= n_gates: zero (b, c) -- a gets the input rows below
    FB_CUDA(cudaMemsetAsync(outs[i] + ng, 0, (m - ng) * sizeof(Fr), st));
    if (ng) launch_spmv(csr, i, w, outs[i], 0, 0, 1, st);
  }
  // bellman appends `input_i * 0 = 0` for every input: a = w_i, b = c = 0
  FB_CUDA(cudaMemcpyAsync(a + ng, w, (size_t)n_in * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
  return FB_OK;
}

// `count` witnesses at once (fb_prove_batch): witness p at w + p * wstride, evaluations of proof p at a/b/c + p * m.
int eval_r1cs_batch(const DevCsr& csr, const Fr* w, uint64_t wstride, uint32_t n_in, Fr* a, Fr* b, Fr* c, uint64_t m,
                    uint32_t count, cudaStream_t st) {
  const uint32_t ng = csr.n_gates;
  Fr* outs[3] = {a, b, c};
  for (int i = 0; i < 3; i++) {
    FB_CUDA(cudaMemset2DAsync(outs[i] + ng, m * sizeof(Fr), 0, (m - ng) * sizeof(Fr), count, st));
    if (ng) launch_spmv(csr, i, w, outs[i], wstride, m, count, st);
  }
  FB_CUDA(cudaMemcpy2DAsync(a + ng, m * sizeof(Fr), w, wstride * sizeof(Fr), (size_t)n_in * sizeof(Fr), count,
                            cudaMemcpyDeviceToDevice, st));
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
    if (ng) launch_spmv(csr, i, w, outs[i], 0, 0, 1, st);
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
