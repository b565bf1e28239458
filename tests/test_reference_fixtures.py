"""Fixtures produced by the REAL reference (tests/golden/ref_*.json, written by bench/ref_node/dump_fixture.js on a
machine with Node + ffjavascript 0.2.59): the oracle (CPU run) and the CUDA path (GPU run) must reproduce them byte for
byte -- commitments, evaluations, key order and every Fiat-Shamir challenge.

This image has no Node, so no such file can be made here; while there is none the tests SKIP with a loud reason (and
DESIGN.md says "parity unpinned").  `bench/ref_node/README.md` is the one-command path that creates them."""
import glob
import json
import os

import pytest

from oracle.py import bn254 as bn, inputs, protocol as pr

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "ref_*.json")))
NO_FIXTURES = ("NO REFERENCE-MADE FIXTURE in tests/golden/ (ref_*.json): parity against genuine ffjavascript output is "
               "UNPINNED on this image (no Node). Run bench/ref_node/dump_fixture.js on a machine with Node and commit "
               "its output.")


def _columns(fx):
    """tests/golden/make_golden.py::columns == bench/ref_node/common.js::buildCase"""
    n = 1 << fx["nbits"]
    cols_f = [inputs.random_column(fx["seed"] * 100 + i, n) for i in range(fx["k"])]
    if fx["selected"] or fx["rotate"]:
        cols_t = [inputs.rotate_right(c) for c in cols_f]
    else:
        perm = inputs.permutation(fx["seed"], n)
        cols_t = [[c[perm[i]] for i in range(n)] for c in cols_f]
    sel_f = sel_t = None
    if fx["selected"]:
        one, zero = bn.fr_to_mont_bytes(1), bytes(32)
        sel_f = one * (n - 1) + zero
        sel_t = zero + one * (n - 1)
    return cols_f, cols_t, sel_f, sel_t


def _check_meta(fx):
    assert fx.get("source") == "reference-node"
    assert fx.get("reference_verifier_accepts", True) is True


def test_js_generator_vectors():
    """bench/ref_node/common.js restates this generator with BigInt; these are the values it must reproduce:
        node -e 'const C=require("./bench/ref_node/common.js");
                 console.log(C.tauFromSeed(1001), C.hex(C.randomColumn(100,2)), C.permutation(31,8))'
    (always runs: it pins the Python side of the contract, and that the vectorised generator of the package agrees)"""
    import numpy as np
    from kzg_grandsums_study_b200 import synthetic
    assert inputs.tau_from_seed(1001) == synthetic.tau_from_seed(1001)
    col = inputs.random_column(100, 2)
    assert bn.fr_vec_to_std_bytes(col) == synthetic.random_fr_std(100, 2).tobytes()
    assert inputs.permutation(31, 8) == [int(x) for x in synthetic.permutation(31, 8)]
    # frozen values (a change of the generator would silently invalidate every committed reference fixture)
    assert inputs.tau_from_seed(1001) == 0x131b5d79e40681d2da44d74b54533efae4d03dd7900d7af197bc95ecdfb81979
    assert np.array_equal(synthetic.permutation(5, 1), np.array([0]))


def _oracle_vs_fixture(fx):
    _check_meta(fx)
    tau = inputs.tau_from_seed(fx["tau_seed"])
    cf, ct, sf, st = _columns(fx)
    fb = [bn.fr_vec_to_std_bytes(c) for c in cf]
    tb = [bn.fr_vec_to_std_bytes(c) for c in ct]
    if fx["nbits"] <= 11:
        prover = pr.grandsum_prover if fx["kind"] == "gs" else pr.grandproduct_prover
        trace = {}
        proof = prover(pr.TrapdoorSrs(tau, fx["ptau_power"]), fb, tb, sf, st, trace=trace)
        assert list(proof["commitments"]) == fx["commitment_keys"]
        assert list(proof["evaluations"]) == fx["evaluation_keys"]
        assert pr.proof_bytes(proof).hex() == fx["proof_bytes"]
        for name, val in fx["challenges"].items():
            assert trace["challenges"][name] == int(val, 16), name
    else:
        from oracle.c import binding as oc
        srs = oc.srs_generate(tau, 2 << fx["nbits"])
        got, ch = oc.prove(fx["kind"], srs, fb, tb, sf, st)
        assert got.hex() == fx["proof_bytes"]
        for name, val in fx["challenges"].items():
            assert bn.fr_from_mont_bytes(ch[name]) == int(val, 16), name


@pytest.mark.skipif(not REF_FIXTURES, reason=NO_FIXTURES)
@pytest.mark.parametrize("path", REF_FIXTURES or ["<none>"], ids=lambda p: os.path.basename(p))
def test_oracle_reproduces_reference_fixture(path):
    _oracle_vs_fixture(json.load(open(path)))


def test_importer_on_a_relabelled_oracle_fixture():
    """the importer itself is exercised even while no reference fixture exists: an oracle-made fixture relabelled as
    reference output must pass, and a single flipped byte must be caught"""
    fx = json.load(open(os.path.join(GOLDEN, "gs_vec_sel_n6_k3.json")))
    fx["source"] = "reference-node"
    _oracle_vs_fixture(fx)
    bad = dict(fx)
    raw = bytearray.fromhex(fx["proof_bytes"])
    raw[200] ^= 1
    bad["proof_bytes"] = raw.hex()
    with pytest.raises(AssertionError):
        _oracle_vs_fixture(bad)


@pytest.mark.gpu
@pytest.mark.skipif(not REF_FIXTURES, reason=NO_FIXTURES)
@pytest.mark.parametrize("path", REF_FIXTURES or ["<none>"], ids=lambda p: os.path.basename(p))
def test_gpu_reproduces_reference_fixture(path, curve, ptau_factory):
    from kzg_grandsums_study_b200.grandsum import mset_eq_kzg_grandsum_prover
    from kzg_grandsums_study_b200.grandproduct import mset_eq_kzg_grandproduct_prover
    from kzg_grandsums_study_b200.polynomial import Evaluations
    fx = json.load(open(path))
    _check_meta(fx)
    cf, ct, sf, st = _columns(fx)
    ev = lambda b: Evaluations(b, curve)
    prover = mset_eq_kzg_grandsum_prover if fx["kind"] == "gs" else mset_eq_kzg_grandproduct_prover
    trace = {}
    proof = prover(ptau_factory(fx["ptau_power"]), [ev(bn.fr_vec_to_std_bytes(c)) for c in cf],
                   [ev(bn.fr_vec_to_std_bytes(c)) for c in ct], ev(sf) if sf else None, ev(st) if st else None, trace=trace)
    assert list(proof["commitments"]) == fx["commitment_keys"]
    assert list(proof["evaluations"]) == fx["evaluation_keys"]
    assert pr.proof_bytes(proof).hex() == fx["proof_bytes"]
    for name, val in fx["challenges"].items():
        assert bn.fr_from_mont_bytes(trace["challenges"][name]) == int(val, 16), name


def test_reference_fixture_status_is_reported():
    """never silent: the run says whether parity is pinned"""
    if not REF_FIXTURES:
        pytest.skip(NO_FIXTURES)
    assert all(json.load(open(p)).get("source") == "reference-node" for p in REF_FIXTURES)
