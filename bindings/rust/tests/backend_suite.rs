//! The reference's shared backend test-suite against the B200 engine — exactly how the reference's own backends are
//! tested (crates/cpu-backend/tests/integration.rs:5, crates/cuda-backend/src/tests.rs:78): every fixture is proved by
//! `B200Engine` and verified by the reference verifier.
use openvm_b200_backend::B200Engine;

openvm_backend_tests::backend_test_suite!(B200Engine);

/// Proofs are byte-identical to the CPU backend's once the proof-of-work witnesses agree: the library returns the smallest
/// witness; the reference CPU prover returns any (rayon `find_any`), so compare everything except the PoW fields.
#[test]
fn proof_equals_cpu_backend_modulo_pow_witnesses() {
    use openvm_stark_backend::StarkEngine;
    use openvm_stark_sdk::{config::baby_bear_poseidon2::BabyBearPoseidon2CpuEngine, test_utils::FibFixture};
    let params = openvm_stark_backend::SystemParams::new_for_testing(16);
    let fib = FibFixture::new(0, 1, 1 << 16);
    let (gpu, cpu) = (B200Engine::new(params.clone()), BabyBearPoseidon2CpuEngine::new(params));
    let (pg, pc) = (fib.prove(&gpu), fib.prove(&cpu));
    assert_eq!(pg.common_main_commit, pc.common_main_commit);
    fib.verify(&gpu, &pg).unwrap();
    fib.verify(&cpu, &pg).unwrap();
}
