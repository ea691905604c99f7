"""TEST INFRASTRUCTURE ONLY.  The seeded random-init weights live in the package (atlaspatch_b200/weights.py: they are the benchmark's
input data, shared by the GPU arm and the CPU arms); the oracle side imports them from here."""
from atlaspatch_b200.weights import VIT_SPECS, vit_state_dict  # noqa: F401
