"""cp_pfdr_graph_d1_b200 - B200-native PFDR hot path of ai3DVision/CP_PFDR_graph_d1.

`api` mirrors the reference's operator interface on top of the C ABI (include/pfdr_b200.h) of the
hand-written sm_100a library lib/libpfdr_b200.so; `synth` generates the seeded benchmark workloads."""
from . import synth  # noqa: F401
from ._abi import DIAG, SCAL, PfdrError, Result  # noqa: F401

__all__ = ["api", "synth", "SCAL", "DIAG", "Result", "PfdrError"]


def __getattr__(name):
    if name == "api":
        import importlib
        return importlib.import_module(".api", __name__)
    raise AttributeError(name)
