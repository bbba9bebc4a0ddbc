/* fawkes_b200.h -- C ABI of the B200-native Groth16 proving backend.
 *
 * Drop-in boundary for fawkes_crypto::backend::bellman_groth16: these entry points
 * are what a Rust FFI crate binds in place of the bellman_ce calls made by the
 * reference (paths relative to the fawkes-crypto repository root):
 *
 *   fb_pk_load / fb_pk_load_circuit
 *       <- bellman::groth16::Parameters::read + WitnessCS::get_gate_iterator
 *          fawkes-crypto/src/backend/bellman_groth16/mod.rs:159-175
 *          fawkes-crypto/src/circuit/r1cs/cs.rs:184-223,248-250
 *   fb_prove
 *       <- bellman::groth16::create_random_proof(bcs, &params.0, rng)
 *          fawkes-crypto/src/backend/bellman_groth16/prover.rs:78-80
 *          (with BellmanCS::synthesize, mod.rs:61-102, folded in: the witness
 *          vectors cross as the raw `Vec<Num<Fr>>` buffers of WitnessCS, cs.rs:99-102)
 *   fb_setup
 *       <- bellman::groth16::generate_random_parameters(bcs, rng)
 *          fawkes-crypto/src/backend/bellman_groth16/setup.rs:17-20
 *   fb_verify
 *       <- bellman::groth16::{prepare_verifying_key, verify_proof}
 *          fawkes-crypto/src/backend/bellman_groth16/verifier.rs:75-81
 *
 * Conventions
 *   - Every field element is a `Num<Fp>` as it sits in memory: 4 x u64 little-endian
 *     limbs in MONTGOMERY form (ff-uint/src/num/mod.rs:21-23); r and s likewise.
 *   - Raw points are `G1Point`/`G2Point` as they sit in memory
 *     (backend/bellman_groth16/group.rs:53-123): G1 = x|y (64 B), G2 = x.c0|x.c1|y.c0|y.c1
 *     (128 B); the point at infinity is all-zero.
 *   - `bellman_params` is the byte string written by bellman's Parameters::write, i.e.
 *     `Parameters.0` of mod.rs:139 (big-endian uncompressed points, u32 BE lengths).
 *   - All functions return 0 on success or a negative FB_ERR_* code; fb_last_error()
 *     gives a thread-local message.  The reference panics (`.unwrap()`, prover.rs:80,
 *     setup.rs:20, verifier.rs:80) or returns io::Error (mod.rs:166-167); the Rust shim
 *     maps codes back to those behaviours.
 *   - There is NO CPU fallback: without a CUDA device every compute entry point
 *     fails with FB_ERR_CUDA.  fb_verify is host code, as in the reference.
 *   - One host thread per fb_ctx; calls are synchronous (prover.rs:63-90 is blocking).
 */
#ifndef FAWKES_B200_H
#define FAWKES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB_OK 0
#define FB_ERR_ARG (-1)
#define FB_ERR_CUDA (-2)
#define FB_ERR_FORMAT (-3)
#define FB_ERR_DOMAIN (-4)   /* bellman SynthesisError::PolynomialDegreeTooLarge */
#define FB_ERR_IDENTITY (-5) /* bellman SynthesisError::UnexpectedIdentity */
#define FB_ERR_DENSITY (-6)  /* query length != structural density of the circuit */
#define FB_ERR_VK (-7)       /* bellman SynthesisError::MalformedVerifyingKey */
#define FB_ERR_HOST (-8)     /* a host-side C++ exception (out of memory, thread creation) stopped at the ABI */

typedef struct fb_ctx fb_ctx;         /* one device + its streams */
typedef struct fb_pk fb_pk;           /* HBM-resident proving key + CSR + workspaces */
typedef struct fb_circuit fb_circuit; /* host-side R1CS in CSR form (parsed gate blob) */

typedef struct fb_pk_info {
  uint32_t n_in;      /* inputs incl. ONE */
  uint32_t n_aux;
  uint32_t n_gates;   /* rows before bellman's input rows */
  uint32_t log_m;     /* domain size 2^log_m */
  uint32_t len_h, len_l, len_a, len_b;
  uint64_t nnz;       /* total non-zeros of A,B,C */
  uint64_t hbm_bytes; /* device memory held by the key */
  uint64_t g1_digit_slots; /* sum over the four G1 MSMs of n x windows = mixed adds per prove (upper bound) */
  uint64_t g2_digit_slots;
  uint32_t msm_window_bits; /* window size chosen for the H MSM */
  uint32_t msm_windows;     /* digits per scalar of the H MSM */
  uint32_t msm_tables;      /* 1 if the key holds the 2^(c w) P window tables */
  uint32_t reserved0;       /* was msm_batch_affine (experiment removed in round 2); always 0 */
  uint64_t table_bytes;     /* part of hbm_bytes held by the window tables */
} fb_pk_info;

/* ---- context --------------------------------------------------------------- */
int fb_init(const int* devices, int ndev, fb_ctx** out); /* ndev must be 1: one process per GPU */
void fb_shutdown(fb_ctx* ctx);
const char* fb_last_error(void);
int fb_device_count(void);
void fb_free(void* p); /* for buffers returned by fb_setup */

/* ---- circuit (parsed gate stream) ------------------------------------------ */
/* gates_brotli: `Parameters.2` (setup.rs:25-32); num_gates: `Parameters.1`. */
int fb_circuit_from_gates(const uint8_t* gates_brotli, size_t len, uint32_t num_gates,
                          uint32_t n_in, uint32_t n_aux, fb_circuit** out);
/* same, from the un-compressed borsh gate stream */
int fb_circuit_from_raw_gates(const uint8_t* gates, size_t len, uint32_t num_gates, uint32_t n_in,
                              uint32_t n_aux, fb_circuit** out);
/* The same parse with the per-term work on the GPU (what WitnessCS::get_gate_iterator + GateStreamedIterator
 * redo on every prove, cs.rs:184-223,248-250: 37-byte terms, "Wrong raw integer" check, canonical -> Montgomery
 * multiply; here once per key, as kernels).  brotli and the walk over the length prefixes stay on the host.
 * Builds the identical circuit (same CSR, same coefficient dictionary in first-appearance order).
 * times_ms: optional float[6] = framing walk, blob upload, kernels, copy back, brotli, device allocation (ms).
 * fb_pk_load uses this path. */
int fb_circuit_from_gates_gpu(fb_ctx* ctx, const uint8_t* gates_brotli, size_t len, uint32_t num_gates,
                              uint32_t n_in, uint32_t n_aux, fb_circuit** out, float* times_ms);
int fb_circuit_from_raw_gates_gpu(fb_ctx* ctx, const uint8_t* gates, size_t len, uint32_t num_gates,
                                  uint32_t n_in, uint32_t n_aux, fb_circuit** out, float* times_ms);
void fb_circuit_free(fb_circuit* c);
int fb_circuit_shape(const fb_circuit* c, uint32_t* n_in, uint32_t* n_aux, uint32_t* n_gates,
                     uint64_t* nnz);

/* ---- proving key ----------------------------------------------------------- */
/* `checked` carries the two booleans of Parameters::read(reader, disallow_points_at_infinity, checked)
 * (mod.rs:159-175; bellman_ce 0.3.5 Parameters::read): bit 0 = checked (every query point must be on its curve
 * and, for G2, in the r-torsion subgroup: [r]P == O), bit 1 = disallow_points_at_infinity.  Always enforced, as by
 * pairing_ce's decoder: coordinates < p, no compression flag, an infinity encoding with every other bit zero;
 * the verifying-key points are always read checked and the ic points may not be at infinity (VerifyingKey::read). */
#define FB_LOAD_CHECKED 1
#define FB_LOAD_NO_INFINITY 2
int fb_pk_load(fb_ctx* ctx, const uint8_t* bellman_params, size_t len, const uint8_t* gates_brotli,
               size_t glen, uint32_t num_gates, int checked, fb_pk** out);
int fb_pk_load_circuit(fb_ctx* ctx, const uint8_t* bellman_params, size_t len,
                       const fb_circuit* circuit, int checked, fb_pk** out);
/* Multi-GPU: this process keeps only base indices [shard, shard+1) * len / nshards of each
 * query (h, l, a, b_g1, b_g2).  fb_prove_partial then returns partial sums. */
int fb_pk_load_shard(fb_ctx* ctx, const uint8_t* bellman_params, size_t len,
                     const fb_circuit* circuit, int checked, int shard, int nshards, fb_pk** out);
void fb_pk_free(fb_pk* pk);
int fb_pk_get_info(const fb_pk* pk, fb_pk_info* info);

/* ---- prove ------------------------------------------------------------------ */
/* inputs[n_in][4] with inputs[0] = ONE (cs.rs:111); aux[n_aux][4]; r, s: Num<Fr>.
 * proof_raw: a.x a.y | b.x.c0 b.x.c1 b.y.c0 b.y.c1 | c.x c.y  (== Proof in memory,
 * prover.rs:13-17).  h_out: optional [m-1][4] H coefficients (natural order), or NULL. */
int fb_prove(fb_ctx* ctx, fb_pk* pk, const uint64_t* inputs, uint32_t n_in, const uint64_t* aux,
             uint32_t n_aux, const uint64_t r[4], const uint64_t s[4], uint8_t proof_raw[256],
             uint64_t* h_out);
/* `count` proofs on one resident key (BASELINE configs[1]: 256 eddsa proofs per run): inputs[i],
 * aux[i] as for fb_prove; r, s: [count][4]; proofs_raw: [count][256].  Keys with a domain of at most 2^16 are proved
 * in chunks of FB_BATCH_P (default 64) proofs with one set of launches per chunk: the proof index is a grid dimension of
 * R1CS evaluation and transforms, and the five MSMs are batched MSMs whose buckets are keyed by (proof, digit).
 * Byte-identical to fb_prove one proof at a time. */
int fb_prove_batch(fb_ctx* ctx, fb_pk* pk, uint32_t count, const uint64_t* const* inputs, uint32_t n_in,
                   const uint64_t* const* aux, uint32_t n_aux, const uint64_t* r, const uint64_t* s,
                   uint8_t* proofs_raw);
/* Streaming proves on one resident key: the reference's prove() alternates witness generation
 * (prover.rs:69-76) and create_random_proof (prover.rs:78-80) on one thread; here submit copies the witness
 * and returns (it blocks only while `depth` proofs are already queued or running; depth <= 0 picks a default),
 * so witness k+1 is generated while proof k runs.  wait blocks until that proof is done and hands out the
 * same bytes fb_prove would; tickets may be collected in any order, each once.  While a stream is open the
 * key must be used through it only.  close drops proofs that have not started (their tickets complete with an
 * error).  Keys with a domain of at most 2^16 prove whatever is queued as ONE batched chunk (see fb_prove_batch), so
 * the batch size follows the caller's submission rate. */
typedef struct fb_stream fb_stream;
int fb_stream_open(fb_ctx* ctx, fb_pk* pk, int depth, fb_stream** out);
int fb_stream_submit(fb_stream* st, const uint64_t* inputs, uint32_t n_in, const uint64_t* aux, uint32_t n_aux,
                     const uint64_t r[4], const uint64_t s[4], uint64_t* ticket);
int fb_stream_wait(fb_stream* st, uint64_t ticket, uint8_t proof_raw[256]);
void fb_stream_close(fb_stream* st);
/* Same, witness already on the device: dev_w = [inputs | aux] as Num<Fr>, 16-byte aligned, READ IN PLACE by the kernels
 * of the prove (no copy): it must stay unchanged until the call returns. */
int fb_prove_device(fb_ctx* ctx, fb_pk* pk, const void* dev_w, const uint64_t r[4],
                    const uint64_t s[4], uint8_t proof_raw[256]);
/* Sharded prove: partial[5] raw affine sums in the order h, l, a, b_g1 (64 B each, slots of
 * 128 B) and b_g2 (128 B): 5 x 128 B.  Combine with fb_prove_finish on any rank. */
int fb_prove_partial(fb_ctx* ctx, fb_pk* pk, const uint64_t* inputs, uint32_t n_in,
                     const uint64_t* aux, uint32_t n_aux, uint8_t partial[640]);
int fb_prove_finish(const uint8_t* bellman_params, size_t len, const uint8_t* partials, int nparts,
                    const uint64_t r[4], const uint64_t s[4], uint8_t proof_raw[256]); /* host only */
/* device-time breakdown of the last fb_prove on this key, milliseconds:
 * [0] h2d  [1] r1cs eval  [2] H pipeline (7 NTTs)  [3] MSMs  [4] d2h+assembly (host)  [5] total */
int fb_prove_timings(const fb_pk* pk, float ms[6]);

/* ---- multi-GPU (one process per GPU) ----------------------------------------------------
 * fb_dist_unique_id on rank 0 (NCCL unique id, 128 bytes), ship it to every rank, then fb_dist_init on
 * each.  A key loaded afterwards with fb_pk_load_shard(shard = rank, nshards = world) also shards the
 * R1CS rows and the H pipeline (four-step NTT, NCCL all-to-all over NVLink) when world is a power of
 * two and the domain is large enough; otherwise those stay replicated and only the MSMs shard.
 * The exchanges of the three evaluation arrays run on a side stream under the transforms of the next array
 * (environment FB_DIST_NO_OVERLAP=1 at fb_dist_init time: one stream, for A/B runs). */
int fb_dist_unique_id(uint8_t id[128]);
int fb_dist_init(fb_ctx* ctx, int rank, int world, const uint8_t id[128]);

/* ---- setup / verify ---------------------------------------------------------- */
/* trapdoor = alpha, beta, gamma, delta, tau (Num<Fr>); generators = the standard BN254
 * ones.  Writes bellman-format Parameters bytes (free with fb_free). */
int fb_setup(fb_ctx* ctx, const fb_circuit* circuit, const uint64_t trapdoor[5][4],
             uint8_t** params_out, size_t* len);
/* Multi-GPU setup: the same byte string, but only the query points that fb_pk_load_shard(shard, nshards) on THIS
 * context keeps are generated (the fixed-base work drops by nshards); every other query point is written as the
 * point at infinity.  The verifying key and ic are complete on every rank.  Load the result with the same
 * shard / nshards on the same context (after fb_dist_init, if the key is to shard its H pipeline too). */
int fb_setup_shard(fb_ctx* ctx, const fb_circuit* circuit, const uint64_t trapdoor[5][4], int shard, int nshards,
                   uint8_t** params_out, size_t* len);
/* vk: alpha_g1 | beta_g2 | gamma_g2 | delta_g2 raw (64+128+128+128 B) then n_ic raw G1 (VK of
 * verifier.rs:12-18 in memory).  inputs: public inputs WITHOUT the leading ONE. */
int fb_verify(const uint8_t* vk_raw, uint32_t n_ic, const uint8_t proof_raw[256],
              const uint64_t* inputs, uint32_t n_inputs, int* ok);

/* ---- benchmark / test harness ------------------------------------------------ */
/* Synthetic random R1CS of SURVEY.md section 8(d).  Witness buffers stay owned by the circuit. */
int fb_circuit_synth(uint64_t n_rows, uint64_t seed, fb_circuit** out);
int fb_circuit_witness(const fb_circuit* c, const uint64_t** inputs, const uint64_t** aux);
int fb_synth_trapdoor(uint64_t seed, uint64_t out[7][4]); /* alpha beta gamma delta tau r s */
/* borrowed views of matrix m (0 A, 1 B, 2 C) in CSR over w = [inputs | aux]; cidx 0 -> ONE,
 * 1 -> -ONE, k >= 2 -> coef_table[k-2] (Montgomery) */
int fb_circuit_csr(const fb_circuit* c, int m, const uint32_t** rowptr, const uint32_t** col,
                   const uint32_t** cidx, uint64_t* nnz, const uint64_t** coef_table, uint64_t* ncoef);
/* instrumentation: kernels launched by the library so far; CUDA-event timing of the dominant
 * kernels on their launching stream (which: 0 G1 bucket accumulation, 1 G2 bucket
 * accumulation, 2 NTT passes, 3 NTT exchanges of a distributed key incl. pack/unpack, 4 digit sorts,
 * 5 bucket reductions) */
uint64_t fb_launch_count(void);
/* 1: run every kernel of a prove on one stream (per-kernel timings are then undisturbed);
 * 0 (default): the witness-only MSMs run on side streams beside the H pipeline */
void fb_set_serial(int on);
/* MSM window tables (2^(c w) P for every window w, W x the base memory, fewer digits per scalar):
 * -1 auto (default: on when they fit in free HBM with headroom), 0 off, 1 on.  Applies to keys loaded
 * afterwards and to fb_test_msm. */
void fb_set_msm_tables(int mode);
/* Keys with a domain of at most 2^16 replay the device side of a prove as one CUDA graph (captured at the
 * first prove of the key): 1 on (default; FB_PROVE_GRAPH=0 in the environment turns it off), 0 off. */
void fb_set_prove_graph(int on);
void fb_kernel_stats_enable(int on);
void fb_kernel_stats_reset(void);
int fb_kernel_stats(int which, uint64_t* launches, double* total_ms);
/* op: 0 mul 1 add 2 sub 3 inv 4 portable-C mul; field: 0 Fr 1 Fq; host buffers [n][4] */
int fb_test_field(fb_ctx* ctx, int field, int op, const uint64_t* a, const uint64_t* b,
                  uint64_t* out, uint64_t n);
/* kind 0 fft 1 ifft 2 coset_fft 3 icoset_fft on 2^log_n host elements, natural order */
int fb_test_ntt(fb_ctx* ctx, int log_n, int kind, uint64_t* data);
/* H coefficients from row evaluations a,b,c (each [2^log_n][4]); out [2^log_n - 1][4] */
int fb_test_h(fb_ctx* ctx, int log_n, const uint64_t* a, const uint64_t* b, const uint64_t* c,
              uint64_t* out, float* ms);
/* distributed H pipeline on ONE GPU: 2^g virtual ranks, exchange by device copies (test of the
 * cyclic/block layouts and kernels without NCCL); out = H coefficients [2^log_n - 1][4] */
int fb_test_dist_h(fb_ctx* ctx, int log_n, int g, const uint64_t* a, const uint64_t* b, const uint64_t* c,
                   uint64_t* out);
/* MSM on host buffers: group 1 (bases 64 B) or 2 (128 B); result raw affine; reps>1 times it */
int fb_test_msm(fb_ctx* ctx, int group, const uint8_t* bases_raw, const uint64_t* scalars,
                uint64_t n, uint8_t* result_raw, int reps, float* ms_per_rep);
/* host only: window bits, digits per scalar and log2(entries per accumulation task) MsmPlan::make picks for n points */
int fb_test_msm_plan(uint32_t n, int table, int* c, int* W, int* task_log);
/* host pairing self-check (no device): addition-chain final exponentiation == plain square-and-multiply, bilinearity
 * e(aG1, bG2) == e(G1, G2)^(ab), multi-pair loop e(P,Q) e(-P,Q) == 1; n rounds from a seeded stream */
int fb_test_pairing(uint64_t seed, int n, int* failures);
/* bases[i] = k_i * G (fixed-base kernel used by setup), raw affine out */
int fb_test_fixed_base(fb_ctx* ctx, int group, const uint64_t* scalars, uint64_t n,
                       uint8_t* out_raw);
/* IMAD-pipe roofline probe: returns 32x32->64 multiply-accumulates per second */
int fb_probe_imad(fb_ctx* ctx, double* mac_per_s);
/* which: 0 carry-chained wide MACs (as issued by the Montgomery rows) per second, 1 PTX Fr
 * multiplies per second, 2 portable-C Fr multiplies per second */
int fb_probe_rate(fb_ctx* ctx, int which, int threads, int blocks_per_sm, double* per_s);
/* Fr multiplies per second (register resident, dependent chains across many warps) */
int fb_probe_fr_mul(fb_ctx* ctx, double* mul_per_s);

#ifdef __cplusplus
}
#endif
#endif /* FAWKES_B200_H */
