"""B200-native execution path for the six LDBC Graphalytics kernels.

Layout: csrc/ (CUDA kernels, C ABI, C++ wrappers and loader), capi.py (ctypes
binding of include/gxb200.h), graphio.py / rmat.py / validator.py (host-side
file formats, synthetic inputs, Graphalytics validation rules).
"""
from . import graphio, rmat, validator  # noqa: F401

__all__ = ["graphio", "rmat", "validator", "capi"]


def __getattr__(name):
    if name == "capi":
        import importlib
        return importlib.import_module(".capi", __name__)
    raise AttributeError(name)
