"""CPU oracle for the VITS hot path — TEST INFRASTRUCTURE ONLY.

Nothing under `oracle/` is part of the product path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import it, and only as the checker (or the timed CPU baseline), never as the
thing shipped: the engine (`piper_b200`) fails loudly when its CUDA library is missing.
"""
