//! Raw bindings to the C ABI of `include/fawkes_b200.h` (the product entry points; the benchmark / self-test
//! helpers at the end of the header are not bound).  Field elements are `Num<Fp>` as they sit in memory:
//! 4 x u64 little-endian Montgomery limbs (ff-uint/src/num/mod.rs:21-23).  Points are `G1Point` / `G2Point` as
//! they sit in memory (backend/bellman_groth16/group.rs:53-123).  Every function returns 0 or a negative
//! `FB_ERR_*` code; `fb_last_error` gives a thread-local message.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const FB_OK: c_int = 0;
pub const FB_ERR_ARG: c_int = -1;
pub const FB_ERR_CUDA: c_int = -2;
pub const FB_ERR_FORMAT: c_int = -3;
pub const FB_ERR_DOMAIN: c_int = -4; // SynthesisError::PolynomialDegreeTooLarge
pub const FB_ERR_IDENTITY: c_int = -5; // SynthesisError::UnexpectedIdentity
pub const FB_ERR_DENSITY: c_int = -6;
pub const FB_ERR_VK: c_int = -7; // SynthesisError::MalformedVerifyingKey
pub const FB_ERR_HOST: c_int = -8; // host-side C++ exception stopped at the ABI (out of memory)

/// bits of the `checked` argument of `fb_pk_load*`: the two booleans of
/// `Parameters::read(reader, disallow_points_at_infinity, checked)` (mod.rs:159)
pub const FB_LOAD_CHECKED: c_int = 1;
pub const FB_LOAD_NO_INFINITY: c_int = 2;

#[repr(C)]
pub struct fb_ctx {
    _p: [u8; 0],
}
#[repr(C)]
pub struct fb_pk {
    _p: [u8; 0],
}
#[repr(C)]
pub struct fb_circuit {
    _p: [u8; 0],
}
#[repr(C)]
pub struct fb_stream {
    _p: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct fb_pk_info {
    pub n_in: u32,
    pub n_aux: u32,
    pub n_gates: u32,
    pub log_m: u32,
    pub len_h: u32,
    pub len_l: u32,
    pub len_a: u32,
    pub len_b: u32,
    pub nnz: u64,
    pub hbm_bytes: u64,
    pub g1_digit_slots: u64,
    pub g2_digit_slots: u64,
    pub msm_window_bits: u32,
    pub msm_windows: u32,
    pub msm_tables: u32,
    pub reserved0: u32,
    pub table_bytes: u64,
}

extern "C" {
    // ---- context
    pub fn fb_init(devices: *const c_int, ndev: c_int, out: *mut *mut fb_ctx) -> c_int;
    pub fn fb_shutdown(ctx: *mut fb_ctx);
    pub fn fb_last_error() -> *const c_char;
    pub fn fb_device_count() -> c_int;
    pub fn fb_free(p: *mut c_void);
    // ---- circuit (Parameters.2 = brotli of the borsh gate stream, setup.rs:25-32; Parameters.1 = num_gates)
    pub fn fb_circuit_from_gates(gates_brotli: *const u8, len: usize, num_gates: u32, n_in: u32, n_aux: u32,
                                 out: *mut *mut fb_circuit) -> c_int;
    pub fn fb_circuit_from_raw_gates(gates: *const u8, len: usize, num_gates: u32, n_in: u32, n_aux: u32,
                                     out: *mut *mut fb_circuit) -> c_int;
    pub fn fb_circuit_from_gates_gpu(ctx: *mut fb_ctx, gates_brotli: *const u8, len: usize, num_gates: u32, n_in: u32,
                                     n_aux: u32, out: *mut *mut fb_circuit, times_ms: *mut f32) -> c_int;
    pub fn fb_circuit_free(c: *mut fb_circuit);
    pub fn fb_circuit_shape(c: *const fb_circuit, n_in: *mut u32, n_aux: *mut u32, n_gates: *mut u32,
                            nnz: *mut u64) -> c_int;
    // ---- proving key
    pub fn fb_pk_load(ctx: *mut fb_ctx, bellman_params: *const u8, len: usize, gates_brotli: *const u8, glen: usize,
                      num_gates: u32, checked: c_int, out: *mut *mut fb_pk) -> c_int;
    pub fn fb_pk_load_circuit(ctx: *mut fb_ctx, bellman_params: *const u8, len: usize, circuit: *const fb_circuit,
                              checked: c_int, out: *mut *mut fb_pk) -> c_int;
    pub fn fb_pk_load_shard(ctx: *mut fb_ctx, bellman_params: *const u8, len: usize, circuit: *const fb_circuit,
                            checked: c_int, shard: c_int, nshards: c_int, out: *mut *mut fb_pk) -> c_int;
    pub fn fb_pk_free(pk: *mut fb_pk);
    pub fn fb_pk_get_info(pk: *const fb_pk, info: *mut fb_pk_info) -> c_int;
    // ---- prove
    pub fn fb_prove(ctx: *mut fb_ctx, pk: *mut fb_pk, inputs: *const u64, n_in: u32, aux: *const u64, n_aux: u32,
                    r: *const u64, s: *const u64, proof_raw: *mut u8, h_out: *mut u64) -> c_int;
    pub fn fb_prove_batch(ctx: *mut fb_ctx, pk: *mut fb_pk, count: u32, inputs: *const *const u64, n_in: u32,
                          aux: *const *const u64, n_aux: u32, r: *const u64, s: *const u64,
                          proofs_raw: *mut u8) -> c_int;
    pub fn fb_prove_device(ctx: *mut fb_ctx, pk: *mut fb_pk, dev_w: *const c_void, r: *const u64, s: *const u64,
                           proof_raw: *mut u8) -> c_int;
    pub fn fb_prove_partial(ctx: *mut fb_ctx, pk: *mut fb_pk, inputs: *const u64, n_in: u32, aux: *const u64,
                            n_aux: u32, partial: *mut u8) -> c_int;
    pub fn fb_prove_finish(bellman_params: *const u8, len: usize, partials: *const u8, nparts: c_int,
                           r: *const u64, s: *const u64, proof_raw: *mut u8) -> c_int;
    pub fn fb_prove_timings(pk: *const fb_pk, ms: *mut f32) -> c_int;
    pub fn fb_stream_open(ctx: *mut fb_ctx, pk: *mut fb_pk, depth: c_int, out: *mut *mut fb_stream) -> c_int;
    pub fn fb_stream_submit(st: *mut fb_stream, inputs: *const u64, n_in: u32, aux: *const u64, n_aux: u32,
                            r: *const u64, s: *const u64, ticket: *mut u64) -> c_int;
    pub fn fb_stream_wait(st: *mut fb_stream, ticket: u64, proof_raw: *mut u8) -> c_int;
    pub fn fb_stream_close(st: *mut fb_stream);
    // ---- multi-GPU (one process per GPU)
    pub fn fb_dist_unique_id(id: *mut u8) -> c_int;
    pub fn fb_dist_init(ctx: *mut fb_ctx, rank: c_int, world: c_int, id: *const u8) -> c_int;
    // ---- setup / verify
    pub fn fb_setup(ctx: *mut fb_ctx, circuit: *const fb_circuit, trapdoor: *const [u64; 4],
                    params_out: *mut *mut u8, len: *mut usize) -> c_int;
    pub fn fb_setup_shard(ctx: *mut fb_ctx, circuit: *const fb_circuit, trapdoor: *const [u64; 4], shard: c_int,
                          nshards: c_int, params_out: *mut *mut u8, len: *mut usize) -> c_int;
    pub fn fb_verify(vk_raw: *const u8, n_ic: u32, proof_raw: *const u8, inputs: *const u64, n_inputs: u32,
                     ok: *mut c_int) -> c_int;
    // ---- policy knobs
    pub fn fb_set_msm_tables(mode: c_int);
    pub fn fb_set_prove_graph(on: c_int);
}
