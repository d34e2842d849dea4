"""Importable alias of the product package `2023-compact-image-compression_b200/` (whose directory name is
not a valid Python identifier).  `import compact_b200 as cb; cb.Encoder(cfg, img).encode()`."""
import importlib as _il
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.abspath(__file__))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
_pkg = _il.import_module("2023-compact-image-compression_b200")
globals().update({k: getattr(_pkg, k) for k in _pkg.__all__})
native = _pkg._native
synth = _il.import_module("2023-compact-image-compression_b200.synth")
codec = _il.import_module("2023-compact-image-compression_b200.codec")
