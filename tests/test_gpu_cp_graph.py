"""SURVEY 8f N3 - connected components, reduced connectivity and the merge step of the cut-pursuit on the device
(csrc/cpgraph.cu) against the oracle restatement (oracle/cp_reduce_oracle_impl.h, pinned to the reference's numbering in
tests/test_oracle_vs_ref.py::test_cp_components_numbering).  Integer results bit-identical; summed weights to rounding."""
import numpy as np
import pytest

import problems
from cp_pfdr_graph_d1_b200 import synth

pytestmark = pytest.mark.gpu


def graph(kind, rng):
    if kind == "grid":
        w = synth.c2_denoise(24, np.float64, diag=False)
        return w["V"], w["Eu"], w["Ev"]
    if kind == "multigraph":
        Eu, Ev = problems.random_graph(3000, 9000, 4)
        return 3000, Eu, Ev
    if kind == "sparse":                                        # many isolated vertices, few edges
        V = 5000
        return V, rng.integers(0, V, 1500).astype(np.int32), rng.integers(0, V, 1500).astype(np.int32)
    if kind == "chain":                                         # one component of diameter V: the worst case for propagation
        V = 200000
        return V, np.arange(V - 1, dtype=np.int32), np.arange(1, V, dtype=np.int32)
    if kind == "big_grid":
        w = synth.c2_denoise(100, np.float32, diag=False)
        return w["V"], w["Eu"], w["Ev"]
    raise KeyError(kind)


@pytest.mark.parametrize("cut", [0.0, 0.3, 0.7, 1.0])
@pytest.mark.parametrize("kind", ["grid", "multigraph", "sparse", "chain", "big_grid"])
def test_components(gpu_api, kind, cut):
    from oracle import oracle as O
    rng = np.random.default_rng(len(kind))
    V, Eu, Ev = graph(kind, rng)
    active = (rng.uniform(0, 1, Eu.shape[0]) < cut).astype(np.uint8)
    Cv, rV = gpu_api.cp_components(V, Eu, Ev, active)
    oCv, orV, _, _ = O.CpGraphOracle().components(V, Eu, Ev, active)
    assert rV == orV and np.array_equal(Cv, oCv)


def test_components_of_a_segmentation(gpu_api):
    """piecewise-constant image: the cut is the set of edges across level sets (what the steepest cut produces)"""
    from oracle import oracle as O
    w = synth.c2_denoise(64, np.float32, diag=False)
    lab = np.round(w["Y"]).astype(int)
    active = (lab[w["Eu"]] != lab[w["Ev"]]).astype(np.uint8)
    Cv, rV = gpu_api.cp_components(w["V"], w["Eu"], w["Ev"], active)
    oCv, orV, _, _ = O.CpGraphOracle().components(w["V"], w["Eu"], w["Ev"], active)
    assert rV == orV and np.array_equal(Cv, oCv)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kind", ["grid", "multigraph", "sparse"])
def test_reduced_edges_and_merge(gpu_api, kind, dtype):
    from oracle import oracle as O
    g = O.CpGraphOracle()
    rng = np.random.default_rng(7)
    V, Eu, Ev = graph(kind, rng)
    E = Eu.shape[0]
    # cut a coarse partition plus a few edges inside its parts (loops in the reduced graph)
    part = rng.integers(0, 12, V) if kind != "grid" else (np.arange(V) * 12 // V)
    active = ((part[Eu] != part[Ev]) | (rng.uniform(0, 1, E) < 0.05)).astype(np.uint8)
    La = rng.uniform(0.1, 1.0, E).astype(dtype)
    La[rng.uniform(0, 1, E) < 0.1] = 0                        # weightless edges are ignored
    Cv, rV = gpu_api.cp_components(V, Eu, Ev, active)
    eps = float(np.finfo(dtype).eps)
    rEu, rEv, rLa = gpu_api.cp_reduced_edges(V, Eu, Ev, active, La, Cv, rV, eps)
    oEu, oEv, oLa = g.reduced_edges(V, Eu, Ev, active, La, Cv, rV, eps)
    assert np.array_equal(rEu, oEu) and np.array_equal(rEv, oEv)
    assert np.allclose(rLa, oLa, rtol=1e-6 if dtype == np.float32 else 1e-13)
    with pytest.raises(gpu_api.PfdrError):
        gpu_api.cp_reduced_edges(V, Eu, Ev, active, La, Cv, rV, eps, capacity=max(1, rEu.shape[0] - 1))
    rX = np.round(rng.uniform(0, 3, rV), 1).astype(dtype)      # many equal values -> many merges
    a1, a0 = active.copy(), active.copy()
    n1 = gpu_api.cp_merge(Eu, Ev, Cv, rX, 1e-3, a1, eps)
    n0 = g.merge(Eu, Ev, Cv, rX, 1e-3, eps, a0)
    assert n1 == n0 and n0 > 0 and np.array_equal(a1, a0)


def test_bad_vertex_number(gpu_api):
    with pytest.raises(gpu_api.PfdrError):
        gpu_api.cp_components(4, np.array([0, 1], np.int32), np.array([1, 7], np.int32), np.zeros(2, np.uint8))
