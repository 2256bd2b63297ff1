"""tikv_b200 — B200-native coprocessor batch-execution engine (TiKV DAG pushdown hot path).

Python here is the test/bench harness above the C ABI (include/b2_copr.h); the product is
tikv_b200/_build/libb2copr.so (CUDA, sm_100a).
"""
from . import ffi, plan  # noqa: F401
