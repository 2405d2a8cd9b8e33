"""TEST INFRASTRUCTURE ONLY.

`oracle/` holds (1) `oracle_np.py`, a numpy restatement of the reference's CPU algorithm
for the sparse-convolution hot path, (2) `build_ref.py`, the recipe that compiles the
reference's own CPU sources into `oracle/_ref/` (git-ignored), and (3) `ref.py`, the loader
for that compiled reference.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
CPU-baseline / `--impl reference` legs may import anything from here; the product package
`minkowskiengine_b200` never does.
"""
