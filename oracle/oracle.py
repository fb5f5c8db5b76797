"""TEST INFRASTRUCTURE ONLY: ctypes access to the two CPU checkers.

  oracle()    -> Binder over oracle/liboracle.so  (serial C restatement, oracle/pfdr_oracle_impl.h)
  reference() -> Binder over oracle/_ref/libref.so (the unmodified reference sources compiled by
                 oracle/build_ref.sh), or None when it has not been built.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from cp_pfdr_graph_d1_b200._abi import Binder  # noqa: E402

ORACLE_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libref.so")


def build(ref=True):
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    if ref:
        subprocess.check_call([os.path.join(_HERE, "build_ref.sh")], stderr=subprocess.DEVNULL)


_o = _r = None


def oracle():
    global _o
    if _o is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        _o = Binder(C.CDLL(ORACLE_SO), "oracle_", returns_status=False, opnorm_seeded=True)
    return _o


def reference():
    global _r
    if _r is None and os.path.exists(REF_SO):
        lib = C.CDLL(REF_SO)
        lib.ref_num_procs.restype = C.c_int
        _r = Binder(lib, "ref_", returns_status=False)
    return _r
