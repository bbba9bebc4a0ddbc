//! Safe wrapper over `fawkes-b200-sys`: the surface of `fawkes_crypto::backend::bellman_groth16`
//! (`setup` / `prover::prove` / `verifier::verify`, `Parameters`, `Proof`) for callers that already hold the raw
//! buffers the reference holds:
//!   * the witness: `WitnessCS.values_input` / `values_aux`, `Vec<Num<Fr>>` (circuit/r1cs/cs.rs:99-102), i.e.
//!     `[[u64; 4]]` Montgomery limbs with `values_input[0] == ONE` (cs.rs:111);
//!   * `Parameters.0` as the bytes bellman's `Parameters::write` produces, `.1` = num_gates, `.2` = the brotli gate
//!     blob (backend/bellman_groth16/mod.rs:139, setup.rs:25-32);
//!   * `Proof` / `VK` as they sit in memory (prover.rs:13-17, verifier.rs:12-18, group.rs:53-123).
//! `backend_patch.rs` shows these calls placed inside the reference's own generic functions.
//!
//! Error behaviour follows the reference: what panics there (`.unwrap()` at prover.rs:80, setup.rs:20,
//! verifier.rs:80; `from_raw_repr(..).unwrap()` at mod.rs:105-120) panics here; what is an `io::Error` there
//! (`Parameters::read`, mod.rs:166-173) is an `io::Error` here.
pub mod osrng;

use fawkes_b200_sys as sys;
use std::ffi::CStr;
use std::io;
use std::os::raw::c_int;
use std::ptr;

pub type Fr = [u64; 4]; // Num<Fr> in memory
pub const PROOF_BYTES: usize = 256; // a.x a.y | b.x.c0 b.x.c1 b.y.c0 b.y.c1 | c.x c.y, raw Montgomery

fn last_error() -> String {
    unsafe {
        let p = sys::fb_last_error();
        if p.is_null() {
            String::new()
        } else {
            CStr::from_ptr(p).to_string_lossy().into_owned()
        }
    }
}

fn check(rc: c_int) {
    if rc != sys::FB_OK {
        // reference: SynthesisError / decoding errors are `.unwrap()`ed -> panic
        panic!("fawkes_b200 error {}: {}", rc, last_error());
    }
}

/// One CUDA device (one process per GPU).  There is no CPU fallback: without a device this panics.
pub struct Context {
    raw: *mut sys::fb_ctx,
}

impl Context {
    pub fn new(device: i32) -> Self {
        let mut raw = ptr::null_mut();
        let dev = [device as c_int];
        check(unsafe { sys::fb_init(dev.as_ptr(), 1, &mut raw) });
        Context { raw }
    }
    /// Join a multi-GPU group (NCCL over NVLink): `id` comes from `dist_unique_id()` on rank 0.
    pub fn dist_init(&mut self, rank: i32, world: i32, id: &[u8; 128]) {
        check(unsafe { sys::fb_dist_init(self.raw, rank, world, id.as_ptr()) });
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { sys::fb_shutdown(self.raw) }
    }
}

pub fn dist_unique_id() -> [u8; 128] {
    let mut id = [0u8; 128];
    check(unsafe { sys::fb_dist_unique_id(id.as_mut_ptr()) });
    id
}

/// The three fields of `Parameters` the prover needs (mod.rs:139): bellman bytes, num_gates, gate blob.
pub struct Parameters {
    pub bellman_bytes: Vec<u8>,
    pub num_gates: u32,
    pub gates_brotli: Vec<u8>,
    pub disallow_points_at_infinity: bool,
    pub checked: bool,
}

/// HBM-resident proving key (bases in Montgomery affine form + window tables, CSR, twiddles, workspaces).
pub struct ProvingKey<'c> {
    ctx: &'c Context,
    raw: *mut sys::fb_pk,
}

impl<'c> ProvingKey<'c> {
    /// `Parameters::read`'s validation happens here, where the points are decoded (mod.rs:159-175): an invalid
    /// key is an `io::Error(InvalidData)`, as in the reference.
    pub fn load(ctx: &'c Context, p: &Parameters) -> io::Result<Self> {
        Self::load_shard(ctx, p, 0, 1)
    }
    /// Multi-GPU: this process keeps base indices `[shard, shard + 1) * len / nshards` of every query.
    pub fn load_shard(ctx: &'c Context, p: &Parameters, shard: i32, nshards: i32) -> io::Result<Self> {
        let mut flags = 0;
        if p.checked {
            flags |= sys::FB_LOAD_CHECKED;
        }
        if p.disallow_points_at_infinity {
            flags |= sys::FB_LOAD_NO_INFINITY;
        }
        // n_in / n_aux are the ic / l query lengths of the bellman byte string (big-endian u32 counts)
        let (n_in, n_aux) = query_lengths(&p.bellman_bytes)?;
        let mut circ = ptr::null_mut();
        let rc = unsafe {
            sys::fb_circuit_from_gates_gpu(ctx.raw, p.gates_brotli.as_ptr(), p.gates_brotli.len(), p.num_gates, n_in,
                                           n_aux, &mut circ, ptr::null_mut())
        };
        if rc != sys::FB_OK {
            return Err(io::Error::new(io::ErrorKind::InvalidData, last_error()));
        }
        let mut raw = ptr::null_mut();
        let rc = unsafe {
            sys::fb_pk_load_shard(ctx.raw, p.bellman_bytes.as_ptr(), p.bellman_bytes.len(), circ, flags, shard, nshards,
                                  &mut raw)
        };
        unsafe { sys::fb_circuit_free(circ) };
        match rc {
            sys::FB_OK => Ok(ProvingKey { ctx, raw }),
            sys::FB_ERR_FORMAT | sys::FB_ERR_DENSITY => Err(io::Error::new(io::ErrorKind::InvalidData, last_error())),
            _ => panic!("fawkes_b200 error {}: {}", rc, last_error()),
        }
    }

    /// `create_proof(circuit, params, r, s)`: deterministic blinding.  On a sharded key (after
    /// `Context::dist_init`) the call is collective: every rank passes the same witness, r and s and every rank
    /// gets the proof.
    pub fn prove_with_rs(&self, values_input: &[Fr], values_aux: &[Fr], r: &Fr, s: &Fr) -> [u8; PROOF_BYTES] {
        let mut proof = [0u8; PROOF_BYTES];
        check(unsafe {
            sys::fb_prove(self.ctx.raw, self.raw, values_input.as_ptr() as *const u64, values_input.len() as u32,
                          values_aux.as_ptr() as *const u64, values_aux.len() as u32, r.as_ptr(), s.as_ptr(),
                          proof.as_mut_ptr(), ptr::null_mut())
        });
        proof
    }

    /// `prove` of prover.rs:63-90: r, s from the OS RNG with the reference's sampling convention (osrng.rs).
    pub fn prove(&self, values_input: &[Fr], values_aux: &[Fr]) -> [u8; PROOF_BYTES] {
        let mut rng = osrng::OsRng::new();
        let (r, s) = (rng.gen_fr(), rng.gen_fr());
        self.prove_with_rs(values_input, values_aux, &r, &s)
    }

    /// Many proofs of one circuit on the resident key (BASELINE configs[1]).
    pub fn prove_batch(&self, witnesses: &[(&[Fr], &[Fr])], rs: &[Fr], ss: &[Fr]) -> Vec<[u8; PROOF_BYTES]> {
        assert!(witnesses.len() == rs.len() && rs.len() == ss.len());
        if witnesses.is_empty() {
            return Vec::new();
        }
        let ins: Vec<*const u64> = witnesses.iter().map(|w| w.0.as_ptr() as *const u64).collect();
        let axs: Vec<*const u64> = witnesses.iter().map(|w| w.1.as_ptr() as *const u64).collect();
        let mut out = vec![[0u8; PROOF_BYTES]; witnesses.len()];
        check(unsafe {
            sys::fb_prove_batch(self.ctx.raw, self.raw, witnesses.len() as u32, ins.as_ptr(), witnesses[0].0.len() as u32,
                                axs.as_ptr(), witnesses[0].1.len() as u32, rs.as_ptr() as *const u64,
                                ss.as_ptr() as *const u64, out.as_mut_ptr() as *mut u8)
        });
        out
    }
}

impl<'c> Drop for ProvingKey<'c> {
    fn drop(&mut self) {
        unsafe { sys::fb_pk_free(self.raw) }
    }
}

/// `setup` of setup.rs:7-35 for an already serialised gate stream: returns bellman `Parameters` bytes.
/// The trapdoor (alpha, beta, gamma, delta, tau) is drawn from the OS RNG like bellman's
/// `generate_random_parameters`; the generators are the standard BN254 ones.
pub fn setup(ctx: &Context, gates_brotli: &[u8], num_gates: u32, n_in: u32, n_aux: u32) -> Vec<u8> {
    let mut rng = osrng::OsRng::new();
    let td: Vec<Fr> = (0..5).map(|_| rng.gen_fr()).collect();
    setup_with_trapdoor(ctx, gates_brotli, num_gates, n_in, n_aux, &td)
}

pub fn setup_with_trapdoor(ctx: &Context, gates_brotli: &[u8], num_gates: u32, n_in: u32, n_aux: u32,
                           trapdoor: &[Fr]) -> Vec<u8> {
    assert!(trapdoor.len() == 5);
    let mut circ = ptr::null_mut();
    check(unsafe {
        sys::fb_circuit_from_gates_gpu(ctx.raw, gates_brotli.as_ptr(), gates_brotli.len(), num_gates, n_in, n_aux,
                                       &mut circ, ptr::null_mut())
    });
    let (mut out, mut len) = (ptr::null_mut(), 0usize);
    let rc = unsafe { sys::fb_setup(ctx.raw, circ, trapdoor.as_ptr(), &mut out, &mut len) };
    unsafe { sys::fb_circuit_free(circ) };
    check(rc);
    let bytes = unsafe { std::slice::from_raw_parts(out, len) }.to_vec();
    unsafe { sys::fb_free(out as *mut _) };
    bytes
}

/// `verify` of verifier.rs:75-81.  `vk_raw` = alpha_g1 | beta_g2 | gamma_g2 | delta_g2 | ic[..] as they sit in
/// memory (verifier.rs:12-18); `inputs` WITHOUT the leading ONE.  Panics where the reference panics: a length
/// mismatch (MalformedVerifyingKey, verifier.rs:80) and malformed encodings (group.rs:53-65, mod.rs:105-120).
pub fn verify(vk_raw: &[u8], proof: &[u8; PROOF_BYTES], inputs: &[Fr]) -> bool {
    assert!(vk_raw.len() >= 448 && (vk_raw.len() - 448) % 64 == 0);
    let n_ic = ((vk_raw.len() - 448) / 64) as u32;
    let mut ok: c_int = 0;
    check(unsafe {
        sys::fb_verify(vk_raw.as_ptr(), n_ic, proof.as_ptr(), inputs.as_ptr() as *const u64, inputs.len() as u32, &mut ok)
    });
    ok != 0
}

fn be32(b: &[u8]) -> u32 {
    u32::from_be_bytes([b[0], b[1], b[2], b[3]])
}

/// (ic length, l length) of a bellman `Parameters` byte string: vk = 64+64+128+128+64+128 bytes, then
/// `u32 BE count | points` sections ic, h, l, a, b_g1, b_g2.
fn query_lengths(p: &[u8]) -> io::Result<(u32, u32)> {
    let short = || io::Error::new(io::ErrorKind::UnexpectedEof, "Parameters truncated");
    if p.len() < 580 {
        return Err(short());
    }
    let n_ic = be32(&p[576..]);
    let mut pos = 580usize + 64 * n_ic as usize;
    if p.len() < pos + 4 {
        return Err(short());
    }
    let n_h = be32(&p[pos..]);
    pos += 4 + 64 * n_h as usize;
    if p.len() < pos + 4 {
        return Err(short());
    }
    Ok((n_ic, be32(&p[pos..])))
}
