"""manta-rs_b200 — B200-native Groth16 proving backend for manta-rs' `ProofSystem::prove` path."""
