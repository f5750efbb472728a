"""ORACLE -- test infrastructure only (see ac_oracle.c).  Parity status:
partially pinned (reference golden vectors + brute-force spec; the Rust
reference cannot run in this image).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package."""
from .binding import Oracle, build_oracle, MATCHKIND_IDS  # noqa: F401
