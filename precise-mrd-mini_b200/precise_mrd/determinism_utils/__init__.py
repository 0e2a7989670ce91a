"""Seeding, environment fingerprint and artifact hashing -- the determinism contract of
``src/precise_mrd/determinism_utils/{seed,hash_artifacts}.py``."""
from .hash_artifacts import hash_dir, hash_file, verify_manifest, write_manifest  # noqa: F401
from .seed import env_fingerprint, set_global_seed  # noqa: F401
