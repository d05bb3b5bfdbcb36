"""Seeded synthetic checkpoints and inputs (no compute): shared by the benchmark, the tests and the oracle tooling."""
