"""Generates tests/golden/*.npz from OUTPUTS OF THE REFERENCE ITSELF (oracle/_ref/libref.so = the unmodified
reference sources compiled by oracle/build_ref.sh).  Run here (where /root/reference exists):

    python tests/golden/make_golden.py

Inputs are the seeded cases of tests/problems.py (regenerated at test time; a checksum of every input array is
stored next to the outputs so that a drifting generator is caught).  The fixtures let the parity tests run on
machines where the reference cannot be built (the GPU box)."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import problems  # noqa: E402
from oracle import oracle as O  # noqa: E402

QUAD = ["grid_identity", "grid_diag_l1_pos", "grid_diag_box2", "grid_diag_recond", "rand_big", "dense_direct",
        "dense_direct_odd", "dense_ata_box"]
SIMPLEX = ["kla", "kla_k16", "kla_recond", "lin", "quad_weighted", "kla_rand_k16", "kla_labels"]


def digest(args):
    """checksum of the inputs, insensitive to last-bit differences: some inputs are derived through BLAS/LAPACK
    (A^tA, lambda_max), whose rounding depends on the host CPU - floats are hashed at ~15 significant bits"""
    h = hashlib.sha256()
    for a in args:
        if isinstance(a, np.ndarray):
            if a.dtype.kind == "f":
                a = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFFF00)
            h.update(np.ascontiguousarray(a).tobytes())
        else:
            h.update(repr(a).encode())
    return h.hexdigest()


def main():
    ref = O.reference()
    assert ref is not None, "build oracle/_ref first (oracle/build_ref.sh)"
    out = {}
    for dt in (np.float32, np.float64):
        tag = np.dtype(dt).name
        for name in QUAD:
            fam, args, kw = problems.quad_case(name, dt)
            r = problems.run_quad(ref, name, dt)
            out[f"quad/{name}/{tag}/X"] = r.X
            out[f"quad/{name}/{tag}/it"] = np.int32(r.it)
            if r.Obj is not None:
                out[f"quad/{name}/{tag}/Obj"] = r.Obj
            out[f"quad/{name}/{tag}/sha"] = np.array(digest(args))
        for name in SIMPLEX:
            args, kw = problems.simplex_case(name, dt)
            r = problems.run_simplex(ref, name, dt)
            out[f"simplex/{name}/{tag}/X"] = r.X
            out[f"simplex/{name}/{tag}/it"] = np.int32(r.it)
            out[f"simplex/{name}/{tag}/Obj"] = r.Obj
            out[f"simplex/{name}/{tag}/sha"] = np.array(digest(args))
        rng = np.random.default_rng(11)
        X = rng.normal(0.3, 1.0, (300, 9)).astype(dt)
        M = rng.uniform(0.2, 1.0, (300, 9)).astype(dt)
        A = rng.uniform(0.5, 2.0, 3).astype(dt)
        out[f"proj/{tag}/X"] = ref.proj_simplex_metric(X.copy(), M, A)
    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    print("wrote", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "reference_outputs.npz")), "bytes")

    # cut-pursuit blocks: what the reference's cut-pursuit returns on the seeded problems of
    # tests/test_oracle_vs_ref.py - components and reduced values after one outer iteration, components after four
    import test_oracle_vs_ref as T
    cp = O.cp_reference()
    cpo = {}
    for mode in ("direct", "ata", "diag"):
        r = T.cp_chain_reference(cp, mode)
        cpo[f"chain/{mode}/Cv"], cpo[f"chain/{mode}/rX"] = r["Cv"].astype(np.int32), r["rX"]
        cpo[f"chain/{mode}/sha"] = np.array(digest([a for a in T.cp_chain_problem(mode) if isinstance(a, np.ndarray)]))
    for kind in ("eeg", "grid", "knn"):
        r = T.cp_numbering_reference(cp, kind)
        cpo[f"numbering/{kind}/Cv"] = r["Cv"].astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "reference_cp_outputs.npz"), **cpo)
    print("wrote", len(cpo), "arrays,", os.path.getsize(os.path.join(HERE, "reference_cp_outputs.npz")), "bytes")


if __name__ == "__main__":
    main()
