// The `cuda` arm of the reference CLI (src/bin/main.rs).  Three edits, everything else in main.rs stays as it is:
//
//   1. imports
//        use tfhe_aes::aes_128::fhe::fhe_impls::cuda_woppbs_1bit::{CudaWoppbs1BitFusedAesEncrypt, CudaWoppbs1BitSboxGalMulPbsAesEncrypt};
//        use tfhe_aes::tfhe::cuda_woppbs_1bit;
//
//   2. two more variants of `enum Implementation` (main.rs:21-27):
//        CudaWoppbs1bit,          // generic circuit, per-SBOX calls coalesced on the GPU
//        CudaWoppbs1bitFused,     // whole rounds on the GPU
//
//   3. two more arms of the `match args.implementation` (main.rs:60-92):

        Implementation::CudaWoppbs1bit => {
            let (client_key, context) = cuda_woppbs_1bit::FheContext::generate_keys_sqrd_lvl_64();
            run_client_server_aes_scenario::<CudaWoppbs1BitSboxGalMulPbsAesEncrypt, _>(&client_key, &context, key, iv, args.number_of_outputs);
        }
        Implementation::CudaWoppbs1bitFused => {
            let (client_key, context) = cuda_woppbs_1bit::FheContext::generate_keys_sqrd_lvl_64();
            run_client_server_aes_scenario::<CudaWoppbs1BitFusedAesEncrypt, _>(&client_key, &context, key, iv, args.number_of_outputs);
        }

// `run_client_server_aes_scenario`, `expand_key` and `encrypt_blocks` (main.rs:97-159) are generic over `Aes128Encrypt`
// and need no change: blocks still fan out with `into_par_iter`, and every worker ends up in `tac_wopbs_coalesced` /
// `tac_aes_encrypt_blocks`, which are safe to call concurrently on one context.
