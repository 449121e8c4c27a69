//! AES-128 on the `cuda_woppbs_1bit` model (B200 GPUs behind `libtfhe_aes_cuda.so`).
//!
//! Two `Aes128Encrypt` implementations:
//! * [`CudaWoppbs1BitSboxGalMulPbsAesEncrypt`] — the generic `fhe_sbox_gal_mul_pbs` circuit, unchanged: every
//!   `Byte::sbox_substitute_and_gal_mul` is one `FheContext::circuit_bootstrap`, i.e. one `tac_wopbs_coalesced` call; the
//!   rayon fan-out over 16 bytes × blocks is merged into batched GPU passes by the library's coalescing queue.
//! * [`CudaWoppbs1BitFusedAesEncrypt`] — the same circuit evaluated round by round on the device with the state
//!   resident in HBM (`tac_aes_encrypt_blocks`, `tac_aes_key_schedule`): no per-SBOX host round trips.
//!
//! Add to `src/aes_128/fhe/fhe_impls/mod.rs`:  `pub mod cuda_woppbs_1bit;`

use crate::aes_128::fhe::data_model::{BitT, Block, Byte, Word};
use crate::aes_128::fhe::{fhe_sbox_gal_mul_pbs, Aes128Encrypt};
use crate::aes_128::{self, SBOX};
use crate::tfhe::cuda_woppbs_1bit::*;
use crate::tfhe::ContextT;

use rayon::iter::ParallelIterator;
use std::array;
use std::sync::{Mutex, OnceLock};
use tfhe_aes_cuda_sys as sys;

impl BitT for BitCt {}

fn context_of(byte: &Byte<BitCt>) -> FheContext {
    byte.0[0].context.clone()
}

impl fhe_sbox_gal_mul_pbs::ByteT for Byte<BitCt> {
    fn bootstrap_assign(&mut self) {
        static IDENTITY_LUT: OnceLock<CudaLut> = OnceLock::new();
        let context = context_of(self);
        let lut = IDENTITY_LUT.get_or_init(|| context.generate_lookup_table(1, 1, |bit| bit as u64));
        self.bits_mut().for_each(|bit| {
            *bit = context.circuit_bootstrap(&[bit], lut).pop().expect("one bit");
        });
    }

    fn sbox_substitute(&self) -> Self {
        static SBOX_LUT: OnceLock<CudaLut> = OnceLock::new();
        let context = context_of(self);
        let lut = SBOX_LUT.get_or_init(|| context.generate_lookup_table(8, 8, |byte| SBOX[byte as usize] as u64));
        let bits: [BitCt; 8] = context.circuit_bootstrap(&self.0.each_ref(), lut).try_into().expect("8 bits");
        Self(bits)
    }

    fn sbox_substitute_and_gal_mul(&self) -> [Self; 3] {
        static SBOX_MUL_LUT: OnceLock<CudaLut> = OnceLock::new();
        let context = context_of(self);
        let lut = SBOX_MUL_LUT.get_or_init(|| {
            context.generate_lookup_table(8, 24, |byte| {
                let s = SBOX[byte as usize];
                ((aes_128::gf_256_mul(s, 1) as u64) << 16) | ((aes_128::gf_256_mul(s, 2) as u64) << 8) | aes_128::gf_256_mul(s, 3) as u64
            })
        });
        let mut bits = context.circuit_bootstrap(&self.0.each_ref(), lut).into_iter();
        array::from_fn(|_| Self(array::from_fn(|_| bits.next().expect("24 bits"))))
    }
}

/// Generic circuit, one GPU call per SBOX (coalesced inside the library)
pub struct CudaWoppbs1BitSboxGalMulPbsAesEncrypt;

impl Aes128Encrypt for CudaWoppbs1BitSboxGalMulPbsAesEncrypt {
    type Ctx = FheContext;

    fn encrypt_block_for_rounds(
        ctx: &Self::Ctx,
        expanded_key: &[Word<<Self::Ctx as ContextT>::Bit>; 44],
        block: Block<<Self::Ctx as ContextT>::Bit>,
        rounds: usize,
    ) -> Block<<Self::Ctx as ContextT>::Bit> {
        fhe_sbox_gal_mul_pbs::encrypt_block_for_rounds(ctx, expanded_key, block, rounds)
    }

    fn key_schedule(
        ctx: &Self::Ctx,
        key_slice: &[Byte<<Self::Ctx as ContextT>::Bit>; 16],
    ) -> [Word<<Self::Ctx as ContextT>::Bit>; 44] {
        fhe_sbox_gal_mul_pbs::key_schedule(ctx, key_slice)
    }
}

/// Whole rounds on the device.  Noise bookkeeping: the library checks the circuit's squared-noise budget statically
/// (`TAC_ERR_NOISE` = the reference's `NoiseTooBig` panic); outputs are tagged like the generic path tags them
/// (last-round SBOX output 8·NOMINAL + round key NOMINAL = 9).
pub struct CudaWoppbs1BitFusedAesEncrypt;

fn flatten<'a>(bytes: impl Iterator<Item = &'a Byte<BitCt>>, out: &mut Vec<u64>) {
    for byte in bytes {
        for bit in &byte.0 {
            out.extend_from_slice(bit.as_words());
        }
    }
}

fn panic_on(ctx: &FheContext, rc: i32) {
    if rc != sys::TAC_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(sys::tac_last_error(ctx.raw())) }.to_string_lossy().into_owned();
        panic!("{msg}");
    }
}

impl Aes128Encrypt for CudaWoppbs1BitFusedAesEncrypt {
    type Ctx = FheContext;

    fn encrypt_block_for_rounds(ctx: &Self::Ctx, expanded_key: &[Word<BitCt>; 44], block: Block<BitCt>, rounds: usize) -> Block<BitCt> {
        let size = ctx.big_lwe_size();
        // the expanded key is uploaded when the caller's array changes (23 MB; main.rs passes the same array for every block)
        static RESIDENT: Mutex<usize> = Mutex::new(0);
        {
            let mut resident = RESIDENT.lock().unwrap();
            if *resident != expanded_key.as_ptr() as usize {
                let mut ks = Vec::with_capacity(44 * 32 * size);
                flatten(expanded_key.iter().flat_map(|w| w.0.iter()), &mut ks);
                panic_on(ctx, unsafe { sys::tac_aes_set_key_schedule(ctx.raw(), ks.as_ptr()) });
                *resident = expanded_key.as_ptr() as usize;
            }
        }
        let mut input = Vec::with_capacity(128 * size);
        flatten(block.iter(), &mut input);
        let mut output = vec![0u64; 128 * size];
        panic_on(ctx, unsafe { sys::tac_aes_encrypt_blocks(ctx.raw(), 1, rounds as i32, 1, input.as_ptr(), output.as_mut_ptr()) });
        let mut cts = output.chunks_exact(size);
        array::from_fn(|_| Byte(array::from_fn(|_| BitCt::with_noise_level(cts.next().unwrap().to_vec(), 9, ctx.clone()))))
    }

    fn key_schedule(ctx: &Self::Ctx, key_slice: &[Byte<BitCt>; 16]) -> [Word<BitCt>; 44] {
        let size = ctx.big_lwe_size();
        let mut key_bits = Vec::with_capacity(128 * size);
        flatten(key_slice.iter(), &mut key_bits);
        let mut ks = vec![0u64; 44 * 32 * size];
        panic_on(ctx, unsafe { sys::tac_aes_key_schedule(ctx.raw(), key_bits.as_ptr(), ks.as_mut_ptr()) });
        let mut cts = ks.chunks_exact(size);
        // words 0..4 are the key itself (fresh, level 1); words 4..44 come out of boot_word (1-bit bootstrap, level 1)
        array::from_fn(|_| Word(array::from_fn(|_| Byte(array::from_fn(|_| BitCt::with_noise_level(cts.next().unwrap().to_vec(), 1, ctx.clone()))))))
    }
}

#[cfg(test)]
mod test {
    use super::*;
    use crate::aes_128::test_helper;
    use std::sync::{Arc, LazyLock};

    static KEYS: LazyLock<(Arc<ClientKey>, FheContext)> = LazyLock::new(|| {
        let (client_key, context) = FheContext::generate_keys_sqrd_lvl_64();
        (client_key.into(), context)
    });

    #[test]
    fn test_light_gal_mul() {
        let (client_key, ctx) = KEYS.clone();
        test_helper::test_light::<CudaWoppbs1BitSboxGalMulPbsAesEncrypt, _>(client_key.as_ref(), &ctx);
    }

    #[test]
    fn test_light_fused() {
        let (client_key, ctx) = KEYS.clone();
        test_helper::test_light::<CudaWoppbs1BitFusedAesEncrypt, _>(client_key.as_ref(), &ctx);
    }

    #[test]
    #[cfg(feature = "long_running_tests")]
    fn test_full_fused() {
        let (client_key, ctx) = KEYS.clone();
        test_helper::test_full::<CudaWoppbs1BitFusedAesEncrypt, _>(client_key.as_ref(), &ctx);
    }
}
