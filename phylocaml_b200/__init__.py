"""phylocaml_b200 -- B200-native tree-scoring engine behind phylocaml's Likelihood /
NonAdditive node-data API. The product is the CUDA library phylocaml_b200/lib/
libphyloc_b200.so (C ABI: include/phylo_engine.h); `engine` is its ctypes binding and
`mlmodel` / `tree` are harness helpers for tests and bench.py."""
from . import engine  # noqa: F401

__all__ = ["engine", "mlmodel", "tree"]
