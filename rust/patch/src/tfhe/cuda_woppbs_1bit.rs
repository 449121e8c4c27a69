//! Model with each ciphertext representing 1 bit, evaluated on NVIDIA B200 GPUs through `libtfhe_aes_cuda.so`.
//!
//! Drop-in sibling of `shortint_woppbs_1bit`: same public surface (`BitCt`, `FheContext`, `ClientKey`, `encode_bit`,
//! `decode_bit`, `generate_lookup_table`, `circuit_bootstrap`, the `generate_keys_sqrd_lvl_*` constructors), same noise
//! bookkeeping and the same panics, so the generic AES code of `aes_128::fhe` runs on it unchanged.  What differs is where
//! the lattice arithmetic happens: the client side (key generation, encryption, decryption) stays on `tfhe-rs`
//! `core_crypto`; the server side (`circuit_bootstrap` = LWE keyswitch + circuit bootstrapping + vertical packing) is one
//! call into the C ABI.  Keys are generated with `core_crypto` directly rather than `shortint::gen_keys` because the
//! device wants the bootstrapping key in the STANDARD domain (the shortint `ServerKey` only keeps the Fourier one).
//!
//! Add to `src/tfhe.rs`:  `pub mod cuda_woppbs_1bit;`
//! Add to `Cargo.toml`:   `tfhe-aes-cuda-sys = { path = "<repo>/rust/tfhe-aes-cuda-sys" }`

use crate::tfhe::engine::ShortintEngine;
use crate::tfhe::{ClientKeyT, ContextT};
use crate::util;

use hashbrown::HashSet;
use std::ffi::CStr;
use std::fmt::{Debug, Formatter};
use std::ops::{BitXor, BitXorAssign};
use std::os::raw::c_int;
use std::sync::atomic::{AtomicU64, Ordering};
use std::sync::{Arc, OnceLock};
use std::time::Instant;
use tfhe::core_crypto::prelude::*;
use tfhe_aes_cuda_sys as sys;
use tracing::debug;

/// Unique id of each non-trivial ciphertext
#[derive(Debug, Clone, Copy, Eq, PartialEq, Hash)]
struct CiphertextId(u64);

/// Squared noise level relative to nominal + the ids of the fresh ciphertexts this one was computed from
#[derive(Clone, Debug)]
pub struct NoiseLevelWithComponents {
    noise_level_squared: u64,
    components: HashSet<CiphertextId>,
}

impl NoiseLevelWithComponents {
    const NOMINAL: u64 = 1;

    fn with_noise_level(noise_level_squared: u64, id: CiphertextId) -> Self {
        Self { noise_level_squared, components: [id].into() }
    }

    fn trivial() -> Self {
        Self { noise_level_squared: 0, components: Default::default() }
    }

    fn add_assign(&mut self, rhs: &Self, max_noise_level_squared: u64) {
        assert!(self.components.is_disjoint(&rhs.components), "noise components not independent");
        self.components.extend(&rhs.components);
        self.noise_level_squared += rhs.noise_level_squared;
        assert!(
            self.noise_level_squared <= max_noise_level_squared,
            "NoiseTooBig: {} > {}",
            self.noise_level_squared,
            max_noise_level_squared
        );
    }
}

/// Ciphertext of a single bit under the big (GLWE) key: `k·N` mask words followed by the body
#[derive(Clone)]
pub struct BitCt {
    ct: Vec<u64>,
    noise_level: NoiseLevelWithComponents,
    pub context: FheContext,
}

impl Debug for BitCt {
    fn fmt(&self, f: &mut Formatter<'_>) -> std::fmt::Result {
        f.debug_struct("BitCt").field("ct", &self.ct).finish()
    }
}

impl BitCt {
    pub fn fresh(ct: Vec<u64>, context: FheContext) -> Self {
        Self::with_noise_level(ct, NoiseLevelWithComponents::NOMINAL, context)
    }

    pub fn with_noise_level(ct: Vec<u64>, noise_level_squared: u64, context: FheContext) -> Self {
        assert_eq!(ct.len(), context.big_lwe_size());
        Self { ct, noise_level: NoiseLevelWithComponents::with_noise_level(noise_level_squared, context.next_ct_id()), context }
    }

    pub fn trivial(bit: Cleartext<u64>, context: FheContext) -> Self {
        let mut ct = vec![0u64; context.big_lwe_size()];
        *ct.last_mut().unwrap() = encode_bit(bit).0;
        Self { ct, noise_level: NoiseLevelWithComponents::trivial(), context }
    }

    pub fn as_words(&self) -> &[u64] {
        &self.ct
    }
}

pub fn encode_bit(bit: Cleartext<u64>) -> Plaintext<u64> {
    assert!(bit.0 < 2, "cleartext out of bounds: {}", bit.0);
    Plaintext(unsafe { sys::tac_encode_bit(bit.0) })
}

pub fn decode_bit(encoding: Plaintext<u64>) -> Cleartext<u64> {
    Cleartext(unsafe { sys::tac_decode_bit(encoding.0) })
}

impl BitXorAssign<&BitCt> for BitCt {
    /// Leveled XOR: wrapping add of the LWE words (2049 of them — host side, like the reference) + noise bookkeeping
    fn bitxor_assign(&mut self, rhs: &Self) {
        for (a, b) in self.ct.iter_mut().zip(&rhs.ct) {
            *a = a.wrapping_add(*b);
        }
        self.noise_level.add_assign(&rhs.noise_level, self.context.inner.params.max_noise_level_squared as u64);
    }
}

impl BitXor for BitCt {
    type Output = Self;

    fn bitxor(mut self, rhs: Self) -> Self::Output {
        self.bitxor_assign(&rhs);
        self
    }
}

/// Lookup table for [`FheContext::circuit_bootstrap`]; registered on the device the first time a context uses it
pub struct CudaLut {
    table: Vec<u64>,
    input_bits: usize,
    output_bits: usize,
    device_id: OnceLock<(usize, c_int)>,
}

struct Device {
    ctx: *mut sys::tac_ctx,
    params: sys::tac_params,
    ct_counter: AtomicU64,
}

// The library locks the context inside every entry point
unsafe impl Send for Device {}
unsafe impl Sync for Device {}

impl Drop for Device {
    fn drop(&mut self) {
        unsafe { sys::tac_ctx_destroy(self.ctx) }
    }
}

#[derive(Clone)]
pub struct FheContext {
    inner: Arc<Device>,
}

fn check(ctx: *mut sys::tac_ctx, rc: c_int) {
    if rc != sys::TAC_OK {
        let msg = unsafe { CStr::from_ptr(sys::tac_last_error(ctx)) }.to_string_lossy().into_owned();
        panic!("tfhe_aes_cuda error {rc}: {msg}");
    }
}

impl FheContext {
    fn next_ct_id(&self) -> CiphertextId {
        CiphertextId(self.inner.ct_counter.fetch_add(1, Ordering::SeqCst))
    }

    pub fn big_lwe_size(&self) -> usize {
        (self.inner.params.glwe_dimension * self.inner.params.polynomial_size) as usize + 1
    }

    pub(crate) fn raw(&self) -> *mut sys::tac_ctx {
        self.inner.ctx
    }

    pub fn generate_keys_sqrd_lvl_1() -> (ClientKey, Self) {
        Self::generate_keys_with_params(1)
    }

    pub fn generate_keys_sqrd_lvl_4() -> (ClientKey, Self) {
        Self::generate_keys_with_params(4)
    }

    pub fn generate_keys_sqrd_lvl_64() -> (ClientKey, Self) {
        Self::generate_keys_with_params(64)
    }

    pub fn generate_keys_sqrd_lvl_256() -> (ClientKey, Self) {
        Self::generate_keys_with_params(256)
    }

    /// Same key material as `shortint::gen_keys` + `WopbsKey::new_wopbs_key_only_for_wopbs` produce for the reference
    /// model, generated through `core_crypto` so that the bootstrapping key is available in the standard domain
    fn generate_keys_with_params(preset: c_int) -> (ClientKey, Self) {
        let mut p = std::mem::MaybeUninit::<sys::tac_params>::uninit();
        assert_eq!(unsafe { sys::tac_params_preset(preset, p.as_mut_ptr()) }, sys::TAC_OK);
        let p = unsafe { p.assume_init() };
        let (glwe_dim, poly, lwe_dim) =
            (GlweDimension(p.glwe_dimension as usize), PolynomialSize(p.polynomial_size as usize), LweDimension(p.lwe_dimension as usize));
        let modulus = CiphertextModulus::new_native();
        let lwe_noise = DynamicDistribution::new_gaussian_from_std_dev(StandardDev(p.lwe_noise_std));
        let glwe_noise = DynamicDistribution::new_gaussian_from_std_dev(StandardDev(p.glwe_noise_std));
        let pfks_noise = DynamicDistribution::new_gaussian_from_std_dev(StandardDev(p.pfks_noise_std));

        // secret keys from an OS-seeded generator (the crate's engine copy, src/tfhe/engine.rs, only carries the encryption generator)
        let mut seeder = new_seeder();
        let mut secret_generator = SecretRandomGenerator::<DefaultRandomGenerator>::new(seeder.seed());
        let glwe_sk: GlweSecretKeyOwned<u64> = allocate_and_generate_new_binary_glwe_secret_key(glwe_dim, poly, &mut secret_generator);
        let lwe_sk: LweSecretKeyOwned<u64> = allocate_and_generate_new_binary_lwe_secret_key(lwe_dim, &mut secret_generator);

        let (glwe_secret_key, lwe_secret_key, bsk, ksk, pfpksk) = ShortintEngine::with_thread_local_mut(|engine| {
            let bsk: LweBootstrapKeyOwned<u64> = par_allocate_and_generate_new_lwe_bootstrap_key(
                &lwe_sk,
                &glwe_sk,
                DecompositionBaseLog(p.pbs_base_log as usize),
                DecompositionLevelCount(p.pbs_level as usize),
                glwe_noise,
                modulus,
                &mut engine.encryption_generator,
            );
            let ksk: LweKeyswitchKeyOwned<u64> = allocate_and_generate_new_lwe_keyswitch_key(
                &glwe_sk.as_lwe_secret_key(),
                &lwe_sk,
                DecompositionBaseLog(p.ks_base_log as usize),
                DecompositionLevelCount(p.ks_level as usize),
                lwe_noise,
                modulus,
                &mut engine.encryption_generator,
            );
            let pfpksk: LwePrivateFunctionalPackingKeyswitchKeyListOwned<u64> =
                par_allocate_and_generate_new_circuit_bootstrap_lwe_pfpksk_list(
                    &glwe_sk.as_lwe_secret_key(),
                    &glwe_sk,
                    DecompositionBaseLog(p.pfks_base_log as usize),
                    DecompositionLevelCount(p.pfks_level as usize),
                    pfks_noise,
                    modulus,
                    &mut engine.encryption_generator,
                );
            (glwe_sk, lwe_sk, bsk, ksk, pfpksk)
        });

        let ctx = unsafe { sys::tac_ctx_create(&p, 0) };
        assert!(!ctx.is_null(), "tac_ctx_create: {}", unsafe { CStr::from_ptr(sys::tac_last_error(std::ptr::null_mut())) }.to_string_lossy());
        // the raw containers are exactly the layouts include/tfhe_aes_cuda.h documents
        check(ctx, unsafe { sys::tac_ctx_upload_keys(ctx, bsk.as_ref().as_ptr(), ksk.as_ref().as_ptr(), pfpksk.as_ref().as_ptr()) });

        let context = FheContext { inner: Arc::new(Device { ctx, params: p, ct_counter: Default::default() }) };
        let client_key = ClientKey { glwe_secret_key, lwe_secret_key, lwe_noise, context: context.clone() };
        (client_key, context)
    }

    /// Same contract as the reference's `generate_lookup_table` (shortint_woppbs_1bit.rs:274-289)
    pub fn generate_lookup_table(&self, input_bits: usize, output_bits: usize, f: impl Fn(u16) -> u64) -> CudaLut {
        assert!(0 < input_bits && input_bits <= 16);
        assert!(0 < output_bits && output_bits <= 64);
        let n = self.inner.params.polynomial_size;
        let f_table: Vec<u64> = (0..1usize << input_bits).map(|v| f(v as u16)).collect();
        let mut table = vec![0u64; output_bits * unsafe { sys::tac_lut_len(input_bits as c_int, n) }];
        let rc = unsafe { sys::tac_generate_lut(input_bits as c_int, output_bits as c_int, n, f_table.as_ptr(), table.as_mut_ptr()) };
        assert_eq!(rc, sys::TAC_OK);
        CudaLut { table, input_bits, output_bits, device_id: OnceLock::new() }
    }

    fn lut_id(&self, lut: &CudaLut) -> c_int {
        let key = Arc::as_ptr(&self.inner) as usize;
        let (owner, id) = *lut.device_id.get_or_init(|| {
            let id = unsafe {
                sys::tac_lut_register(self.raw(), lut.input_bits as c_int, lut.output_bits as c_int, lut.table.as_ptr(), lut.table.len())
            };
            assert!(id >= 0, "tac_lut_register failed");
            (key, id)
        });
        assert_eq!(owner, key, "a CudaLut is bound to the first context that used it (one parameter set per process, like the reference's OnceLock LUTs)");
        id
    }

    /// Circuit bootstrap with the given bits as input.  Callers arrive from rayon workers (16 SBOX bytes × blocks);
    /// `tac_wopbs_coalesced` merges the concurrent calls into one batched pass on the device.
    pub fn circuit_bootstrap(&self, bits: &[&BitCt], lut: &CudaLut) -> Vec<BitCt> {
        assert_eq!(bits.len(), lut.input_bits);
        let start = Instant::now();
        let size = self.big_lwe_size();
        let mut input = Vec::with_capacity(bits.len() * size);
        for bit in bits {
            input.extend_from_slice(&bit.ct);
        }
        let mut output = vec![0u64; lut.output_bits * size];
        check(self.raw(), unsafe { sys::tac_wopbs_coalesced(self.raw(), self.lut_id(lut), 1, input.as_ptr(), output.as_mut_ptr()) });
        // Lemma 3.2 of eprint 2017/430: the number of selector inputs multiplies the error variance
        let output_noise_level_squared = NoiseLevelWithComponents::NOMINAL * bits.len() as u64;
        let bit_cts =
            output.chunks_exact(size).map(|ct| BitCt::with_noise_level(ct.to_vec(), output_noise_level_squared, self.clone())).collect();
        debug!("multivalued circuit bootstrap {:?}", start.elapsed());
        bit_cts
    }
}

impl ContextT for FheContext {
    type Bit = BitCt;

    fn trivial(&self, bit: Cleartext<u64>) -> BitCt {
        BitCt::trivial(bit, self.clone())
    }
}

pub struct ClientKey {
    glwe_secret_key: GlweSecretKeyOwned<u64>,
    #[allow(unused)]
    lwe_secret_key: LweSecretKeyOwned<u64>,
    lwe_noise: DynamicDistribution<u64>,
    context: FheContext,
}

impl ClientKeyT for ClientKey {
    type Bit = BitCt;

    fn encrypt(&self, bit: Cleartext<u64>) -> BitCt {
        let ct = ShortintEngine::with_thread_local_mut(|engine| {
            allocate_and_encrypt_new_lwe_ciphertext(
                &self.glwe_secret_key.as_lwe_secret_key(),
                encode_bit(bit),
                self.lwe_noise,
                CiphertextModulus::new_native(),
                &mut engine.encryption_generator,
            )
        });
        BitCt::fresh(ct.into_container(), self.context.clone())
    }

    fn decrypt(&self, bit: &BitCt) -> Cleartext<u64> {
        let ct = LweCiphertext::from_container(bit.ct.as_slice(), CiphertextModulus::new_native());
        decode_bit(decrypt_lwe_ciphertext(&self.glwe_secret_key.as_lwe_secret_key(), &ct))
    }
}

impl ClientKey {
    /// Secret and evaluation keys as a wire file (csrc/wire.cpp) — e.g. to run the Python / C++ hosts of the B200
    /// repository on exactly the keys of this process
    pub fn secret_key_words(&self) -> (&[u64], &[u64]) {
        (self.glwe_secret_key.as_ref(), self.lwe_secret_key.as_ref())
    }
}

#[allow(unused)]
fn bits_of(val: u16) -> [u8; 16] {
    util::u16_to_bits(val)
}
