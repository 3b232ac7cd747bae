"""CPU oracle of the ParELAGMC per-sample hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import
this package; the product (`parelagmc_b200/`) never does.  See `oracle/pmc_oracle.h` for scope and
parity status ("parity unpinned" for the third-party pieces: TRNG, ParELAG, MFEM, hypre are not vendored
under /root/reference and cannot be built here).
"""
