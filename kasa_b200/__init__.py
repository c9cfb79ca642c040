"""kasa_b200 — B200-native (sm_100a) implementation of kASA's `identify` hot path and `build` index-sort path.

The product is libkasa_b200.so (hand-written CUDA behind the C ABI of include/kasa_b200.h); this package is the
thin Python host mirror used by the tests and bench.py.  There is no CPU fallback: importing works anywhere,
but every compute call needs the built library and a CUDA device and fails loudly otherwise.
"""
from .capi import Context, KasaError, Params, library_path, load_library  # noqa: F401

__version__ = "0.1.0"
