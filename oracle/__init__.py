"""TEST INFRASTRUCTURE -- CPU restatements of the reference algorithms.

Nothing under `oracle/` is product code: only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.
"""
