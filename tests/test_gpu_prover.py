"""GPU: whole proofs through the drop-in entry points, byte-identical to the CPU oracle's restatement of the
reference provers, accepted by the oracle's restatement of the reference verifiers."""
import json
import os

import pytest

from oracle.py import bn254 as bn, inputs, protocol as pr

pytestmark = pytest.mark.gpu
R = bn.R
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_case(curve, seed, nbits, k=1, selected=False, rotate=False):
    """inputs as the reference tests build them (test/mset_eq_kzg_grandsum.test.js:24-104) from seeded PRNGs"""
    from kzg_grandsums_study_b200.polynomial import Evaluations
    n = 1 << nbits
    cols_f = [inputs.random_column(seed * 100 + i, n) for i in range(k)]
    sel_f = sel_t = None
    if selected or rotate:
        cols_t = [inputs.rotate_right(c) for c in cols_f]
    else:
        perm = inputs.permutation(seed, n)
        cols_t = [[c[perm[i]] for i in range(n)] for c in cols_f]
    if selected:
        one, zero = bn.fr_to_mont_bytes(1), bytes(32)
        sel_f = one * (n - 1) + zero          # selF[n-1] = 0   (:68-71)
        sel_t = zero + one * (n - 1)          # selT[0] = 0
    std = [bn.fr_vec_to_std_bytes(c) for c in cols_f], [bn.fr_vec_to_std_bytes(c) for c in cols_t]
    ev = lambda b: Evaluations(b, curve)
    dev_args = ([ev(b) for b in std[0]], [ev(b) for b in std[1]],
                ev(sel_f) if selected else None, ev(sel_t) if selected else None)
    return std, (sel_f, sel_t), dev_args


def run_both(kind, curve, tau, ptau_factory, seed, nbits, k=1, selected=False, rotate=False, ptau_power=None):
    from kzg_grandsums_study_b200.grandsum import mset_eq_kzg_grandsum_prover
    from kzg_grandsums_study_b200.grandproduct import mset_eq_kzg_grandproduct_prover
    std, sels, dev_args = make_case(curve, seed, nbits, k, selected, rotate)
    power = ptau_power if ptau_power is not None else nbits
    path = ptau_factory(power)
    gpu_prover = mset_eq_kzg_grandsum_prover if kind == "gs" else mset_eq_kzg_grandproduct_prover
    cpu_prover = pr.grandsum_prover if kind == "gs" else pr.grandproduct_prover
    trace = {}
    fs, ts = dev_args[0], dev_args[1]
    if k == 1:
        fs, ts = fs[0], ts[0]                 # a single Evaluations instead of a list (prover.js:22-27)
    got = gpu_prover(path, fs, ts, dev_args[2], dev_args[3], trace=trace)
    otrace = {}
    want = cpu_prover(pr.TrapdoorSrs(tau, power), std[0], std[1], sels[0], sels[1], trace=otrace)
    run_both.last_path = path
    return got, want, trace, otrace


def package_verify(kind, proof, nbits):
    """the drop-in verifier of the package (host code, real pairing against [tau]_2 of the same .ptau)"""
    from kzg_grandsums_study_b200.grandsum import mset_eq_kzg_grandsum_verifier
    from kzg_grandsums_study_b200.grandproduct import mset_eq_kzg_grandproduct_verifier
    v = mset_eq_kzg_grandsum_verifier if kind == "gs" else mset_eq_kzg_grandproduct_verifier
    return v(run_both.last_path, proof, nbits)


def assert_same_proof(got, want, trace, otrace):
    assert list(got["commitments"]) == list(want["commitments"])
    assert list(got["evaluations"]) == list(want["evaluations"])
    for name in want["commitments"]:
        assert got["commitments"][name] == want["commitments"][name], name
    for name in want["evaluations"]:
        assert got["evaluations"][name] == want["evaluations"][name], name
    for name, val in otrace["challenges"].items():
        assert bn.fr_from_mont_bytes(trace["challenges"][name]) == val, name
    assert pr.proof_bytes(got) == pr.proof_bytes(want)


def test_appendix_f_vector(curve, tmp_path):
    """SURVEY.md Appendix F through the GPU path, ptau file written by the oracle"""
    from kzg_grandsums_study_b200.grandsum import mset_eq_kzg_grandsum_prover
    from kzg_grandsums_study_b200.polynomial import Evaluations
    from oracle.py import ptau as opt
    tau = 0x1234567
    path = str(tmp_path / "f.ptau")
    opt.write_ptau(path, 2, tau)
    F = Evaluations(bn.fr_vec_to_std_bytes([1, 2, 3, 4]), curve)
    T = Evaluations(bn.fr_vec_to_std_bytes([4, 1, 2, 3]), curve)
    proof = mset_eq_kzg_grandsum_prover(path, F, T)
    assert proof["commitments"]["F"].hex() == (
        "cd055c2b428e58a495972339985546dba726573ea29b6270ade6ec357fbaa204"
        "a504b976f13a4e456864b995f8e82c337678c53710968db0930eb9f60323dc25")
    assert proof["evaluations"]["fxi"].hex() == "6228f33f6983e1f743118a275b003d1617857745952bd101192d11346285be1c"
    assert bn.g1_from_bytes(proof["commitments"]["Wxiw"]) == (
        10319162343374394764343463619286087936657639163999839545968593063611593848571,
        21696086493473458144269509860650732329214412569332322931926248127506125646569)
    assert bn.fr_from_mont_bytes(proof["evaluations"]["sxiw"]) == \
        3812625113219582593042459501361184331188234103107494215927524191815485429995
    assert pr.grandsum_verifier(proof, 2, tau=tau)


@pytest.mark.parametrize("kind", ["gs", "gp"])
@pytest.mark.parametrize("nbits", [1, 2, 3, 5, 8, 10])
def test_plain_proofs(kind, nbits, curve, tau, ptau_factory):
    """the reference's plain test shape: T = F rotated right by one (test/...test.js:24-37), nBits in [1, 10]"""
    got, want, tr, otr = run_both(kind, curve, tau, ptau_factory, seed=nbits, nbits=nbits, rotate=True)
    assert_same_proof(got, want, tr, otr)
    verifier = pr.grandsum_verifier if kind == "gs" else pr.grandproduct_verifier
    assert verifier(got, nbits, tau=tau)
    assert package_verify(kind, got, nbits) is True


@pytest.mark.parametrize("kind", ["gs", "gp"])
@pytest.mark.parametrize("nbits,k,selected", [(4, 3, False), (6, 10, False), (4, 1, True), (7, 1, True),
                                              (5, 4, True), (3, 2, True)])
def test_vector_and_selected_proofs(kind, nbits, k, selected, curve, tau, ptau_factory):
    """vector (:39-57), selected (:59-78) and selected-vector (:80-104) shapes"""
    got, want, tr, otr = run_both(kind, curve, tau, ptau_factory, seed=7 + nbits, nbits=nbits, k=k, selected=selected)
    assert_same_proof(got, want, tr, otr)
    verifier = pr.grandsum_verifier if kind == "gs" else pr.grandproduct_verifier
    assert verifier(got, nbits, tau=tau)
    assert package_verify(kind, got, nbits) is True
    got["evaluations"][next(iter(got["evaluations"]))] = bn.fr_to_mont_bytes(12345)
    assert package_verify(kind, got, nbits) is False


@pytest.mark.parametrize("kind", ["gs", "gp"])
def test_config_c1_c2(kind, curve, tau, ptau_factory):
    """BASELINE configs: C1 nBits=8 on a power-11 ptau; C2 nBits=11, grand-sum vs grand-product on the same inputs"""
    got, want, tr, otr = run_both(kind, curve, tau, ptau_factory, seed=1, nbits=8, rotate=True, ptau_power=11)
    assert_same_proof(got, want, tr, otr)
    got, want, tr, otr = run_both(kind, curve, tau, ptau_factory, seed=2, nbits=11, rotate=True, ptau_power=11)
    assert_same_proof(got, want, tr, otr)


def test_golden_fixtures(curve, ptau_factory):
    """committed golden proofs (tests/golden/*.json, generated by tests/golden/make_golden.py from the oracle)"""
    from kzg_grandsums_study_b200.grandsum import mset_eq_kzg_grandsum_prover
    from kzg_grandsums_study_b200.grandproduct import mset_eq_kzg_grandproduct_prover
    names = sorted(f for f in os.listdir(GOLDEN) if f.endswith(".json"))
    assert names
    for name in names:
        g = json.load(open(os.path.join(GOLDEN, name)))
        assert g["tau_seed"] == 1001
        std, sels, dev_args = make_case(curve, g["seed"], g["nbits"], g["k"], g["selected"], g["rotate"])
        prover = mset_eq_kzg_grandsum_prover if g["kind"] == "gs" else mset_eq_kzg_grandproduct_prover
        proof = prover(ptau_factory(g["ptau_power"]), dev_args[0], dev_args[1], dev_args[2], dev_args[3])
        assert pr.proof_bytes(proof).hex() == g["proof_bytes"], name
        assert list(proof["commitments"]) == g["commitment_keys"] and list(proof["evaluations"]) == g["evaluation_keys"]


def test_prover_errors(curve, tau, ptau_factory):
    """argument checks and protocol failures raise the reference's messages"""
    from kzg_grandsums_study_b200 import KzgError
    from kzg_grandsums_study_b200.grandsum import mset_eq_kzg_grandsum_prover as gs
    from kzg_grandsums_study_b200.grandproduct import mset_eq_kzg_grandproduct_prover as gp
    from kzg_grandsums_study_b200.polynomial import Evaluations
    path = ptau_factory(3)
    ev = lambda v: Evaluations(bn.fr_vec_to_std_bytes(v), curve)
    f = inputs.random_column(1, 8)
    t = inputs.rotate_right(f)
    with pytest.raises(ValueError, match="lengths of the two vector multisets"):
        gs(path, [ev(f), ev(f)], [ev(t)])
    with pytest.raises(ValueError, match="greater than 0"):
        gs(path, [], [])
    with pytest.raises(ValueError, match="0-th multiset buffers must have the same length"):
        gs(path, ev(f), ev(t[:4]))
    with pytest.raises(ValueError, match="must all have the same length"):
        gs(path, [ev(f), ev(f[:4])], [ev(t), ev(t[:4])])
    with pytest.raises(ValueError, match="power of two"):
        gs(path, ev(f[:6]), ev(t[:6]))
    with pytest.raises(ValueError, match="selection buffers must have the same length\\."):
        gs(path, ev(f), ev(t), Evaluations.getOneEvals(8, curve), Evaluations.getOneEvals(4, curve))
    with pytest.raises(ValueError, match="same length as the multiset buffers"):
        gs(path, ev(f), ev(t), Evaluations.getOneEvals(4, curve), Evaluations.getOneEvals(4, curve))
    f16 = inputs.random_column(2, 16)
    with pytest.raises(ValueError, match="not sufficiently large"):
        gs(path, ev(f16), ev(inputs.rotate_right(f16)))
    bad = list(t)
    bad[2] = (bad[2] + 1) % R
    with pytest.raises(KzgError, match="The grand-sum polynomial S is not well calculated"):
        gs(path, ev(f), ev(bad))
    with pytest.raises(KzgError, match="The grand-product polynomial Z is not well calculated"):
        gp(path, ev(f), ev(bad))
    # non-binary selectors: S / Z are still "well calculated" only if the selected multisets agree; a selector value
    # of 2 on both sides of the same row keeps the sums equal but breaks selF - selF^2 = 0 -> not divisible
    two = bn.fr_to_mont_bytes(2)
    one = bn.fr_to_mont_bytes(1)
    sel = Evaluations(two + one * 7, curve)
    with pytest.raises(KzgError, match="Polynomial is not divisible"):
        gs(path, ev(f), ev(f), sel, Evaluations(two + one * 7, curve))


@pytest.mark.parametrize("nbits", [16])
def test_config_c3_selected_vector(nbits, curve, tau, ptau_factory):
    """BASELINE C3: selected-vector grand-sum at n = 2^16, k = 4 (oracle with closed-form commitments)"""
    got, want, tr, otr = run_both("gs", curve, tau, ptau_factory, seed=3, nbits=nbits, k=4, selected=True)
    assert_same_proof(got, want, tr, otr)
    assert pr.grandsum_verifier(got, nbits, tau=tau)


def test_column_buffer_types(curve, tau, ptau_factory):
    """host columns may be bytes, bytearray, numpy arrays or (pinned) torch tensors -- same proof"""
    import numpy as np
    import torch
    from kzg_grandsums_study_b200.grandsum import mset_eq_kzg_grandsum_prover
    from kzg_grandsums_study_b200.polynomial import Evaluations
    nbits = 6
    f = inputs.random_column(5, 1 << nbits)
    fb, tb = bn.fr_vec_to_std_bytes(f), bn.fr_vec_to_std_bytes(inputs.rotate_right(f))
    path = ptau_factory(nbits)
    ref = pr.proof_bytes(mset_eq_kzg_grandsum_prover(path, Evaluations(fb, curve), Evaluations(tb, curve)))
    variants = [
        (bytearray(fb), bytearray(tb)),
        (np.frombuffer(fb, dtype=np.uint64).reshape(-1, 4).copy(), np.frombuffer(tb, dtype=np.uint8).copy()),
        (torch.frombuffer(bytearray(fb), dtype=torch.uint8).pin_memory(), torch.frombuffer(bytearray(tb), dtype=torch.uint8).pin_memory()),
    ]
    for a, b in variants:
        ea, eb = Evaluations(a, curve), Evaluations(b, curve)
        assert ea.length() == 1 << nbits and ea.getEvaluation(3) == fb[96:128]
        assert pr.proof_bytes(mset_eq_kzg_grandsum_prover(path, ea, eb)) == ref


@pytest.mark.parametrize("nbits", [20, 22])
def test_config_c4_full_size_properties(nbits, curve, tau, ptau_factory):
    """BASELINE C4 (n = 2^20, 2^22; T = PRNG permutation of F): sizes the Python oracle cannot reach.  Size-independent
    properties: the proof is accepted by the drop-in verifier (real pairing against [tau]_2 of the same .ptau) and by
    the oracle's trapdoor check; [F] and [T] equal the closed forms F(tau) G1 / T(tau) G1 computed through a different
    device path (Horner evaluation + fixed-base product); a tampered proof is rejected; the proof is reproducible."""
    import numpy as np
    from kzg_grandsums_study_b200 import synthetic
    from kzg_grandsums_study_b200.grandsum import mset_eq_kzg_grandsum_prover, mset_eq_kzg_grandsum_verifier
    from kzg_grandsums_study_b200.polynomial import Evaluations, Polynomial
    n = 1 << nbits
    f = synthetic.random_fr_std(4 if nbits == 20 else 5, n)
    t = np.ascontiguousarray(f[synthetic.permutation(4 if nbits == 20 else 5, n)])
    path = ptau_factory(nbits)
    proof = mset_eq_kzg_grandsum_prover(path, Evaluations(f, curve), Evaluations(t, curve))
    assert list(proof["commitments"]) == ["F", "T", "S", "Q", "Wxi", "Wxiw"]
    assert mset_eq_kzg_grandsum_verifier(path, proof, nbits) is True
    assert pr.grandsum_verifier(proof, nbits, tau=tau)
    # closed forms: commit(p) = p(tau) G1 with p = iNTT(batchToMontgomery(column))
    for name, col in (("F", f), ("T", t)):
        p = Polynomial.fromEvaluations(curve.Fr.batchToMontgomery(col.tobytes()), curve)
        p_tau = bn.fr_from_mont_bytes(p.evaluate(bn.fr_to_mont_bytes(tau)))
        assert proof["commitments"][name] == bn.g1_to_bytes(bn.g1_mul_gen(p_tau)), name
    again = mset_eq_kzg_grandsum_prover(path, Evaluations(f, curve), Evaluations(t, curve))
    assert pr.proof_bytes(again) == pr.proof_bytes(proof)
    bad = {"commitments": dict(proof["commitments"]), "evaluations": dict(proof["evaluations"])}
    bad["evaluations"]["sxiw"] = bn.fr_to_mont_bytes(7)
    assert mset_eq_kzg_grandsum_verifier(path, bad, nbits) is False
    # multisets that differ in one element are refused with the reference's message
    t2 = t.copy()
    t2[12345, 0] ^= np.uint64(1)
    from kzg_grandsums_study_b200 import KzgError
    with pytest.raises(KzgError, match="The grand-sum polynomial S is not well calculated"):
        mset_eq_kzg_grandsum_prover(path, Evaluations(f, curve), Evaluations(t2, curve))


def _prove_pair(kind, curve, tau, ptau_factory, nbits, cols_f, cols_t, sel_f=None, sel_t=None):
    """GPU proof and oracle proof for explicit integer columns (selectors: lists of 0/1 or None)"""
    from kzg_grandsums_study_b200.grandsum import mset_eq_kzg_grandsum_prover
    from kzg_grandsums_study_b200.grandproduct import mset_eq_kzg_grandproduct_prover
    from kzg_grandsums_study_b200.polynomial import Evaluations
    path = ptau_factory(nbits)
    fb = [bn.fr_vec_to_std_bytes(c) for c in cols_f]
    tb = [bn.fr_vec_to_std_bytes(c) for c in cols_t]
    sfb = bn.fr_vec_to_mont_bytes(sel_f) if sel_f is not None else None
    stb = bn.fr_vec_to_mont_bytes(sel_t) if sel_t is not None else None
    ev = lambda b: Evaluations(b, curve)
    gpu = mset_eq_kzg_grandsum_prover if kind == "gs" else mset_eq_kzg_grandproduct_prover
    cpu = pr.grandsum_prover if kind == "gs" else pr.grandproduct_prover
    got = gpu(path, [ev(b) for b in fb], [ev(b) for b in tb], ev(sfb) if sfb else None, ev(stb) if stb else None)
    want = cpu(pr.TrapdoorSrs(tau, nbits), fb, tb, sfb, stb)
    run_both.last_path = path
    return got, want


@pytest.mark.parametrize("kind", ["gs", "gp"])
def test_randomised_shapes(kind, curve, tau, ptau_factory):
    """a seeded sweep over (nBits, k, selectors, selector density): GPU proof == oracle proof, byte for byte"""
    import random
    rng = random.Random(20261018 if kind == "gs" else 7)
    for case in range(24):
        nbits = rng.randint(1, 9)
        n = 1 << nbits
        k = rng.choice([1, 1, 2, 3, 5])
        selected = rng.random() < 0.5 and n >= 4
        cols_f = [inputs.random_column(rng.randrange(1 << 30), n) for _ in range(k)]
        if selected:
            # binary selectors with the same number of ones on both sides; selected rows of T are a permutation of the
            # selected rows of F (row-wise across the k columns), unselected rows are unrelated
            ones = rng.randint(1, n - 1)
            rows_f = sorted(rng.sample(range(n), ones))
            rows_t = sorted(rng.sample(range(n), ones))
            perm = rows_f[:]
            rng.shuffle(perm)
            cols_t = [inputs.random_column(rng.randrange(1 << 30), n) for _ in range(k)]
            for dst, src in zip(rows_t, perm):
                for c in range(k):
                    cols_t[c][dst] = cols_f[c][src]
            sel_f = [1 if i in set(rows_f) else 0 for i in range(n)]
            sel_t = [1 if i in set(rows_t) else 0 for i in range(n)]
        else:
            perm = list(range(n))
            rng.shuffle(perm)
            cols_t = [[c[perm[i]] for i in range(n)] for c in cols_f]
            sel_f = sel_t = None
        got, want = _prove_pair(kind, curve, tau, ptau_factory, nbits, cols_f, cols_t, sel_f, sel_t)
        assert pr.proof_bytes(got) == pr.proof_bytes(want), (case, nbits, k, selected)
        assert list(got["commitments"]) == list(want["commitments"])
    assert package_verify(kind, got, nbits) is True


@pytest.mark.parametrize("kind", ["gs", "gp"])
def test_degenerate_columns(kind, curve, tau, ptau_factory):
    """inputs outside the reference's comfort zone where the mathematically correct result is still
    well defined: all-zero columns (every commitment is the point at infinity), constant columns (degree-0 witness
    polynomials), repeated values, small values"""
    n = 16
    for cols in ([0] * n, [7] * n, [1, 2] * (n // 2), list(range(n)), [R - 1] * n, [0] * (n - 1) + [5]):
        f = list(cols)
        t = f[3:] + f[:3]
        got, want = _prove_pair(kind, curve, tau, ptau_factory, 4, [f], [t])
        assert pr.proof_bytes(got) == pr.proof_bytes(want), cols[:4]
    zero = _prove_pair(kind, curve, tau, ptau_factory, 4, [[0] * n], [[0] * n])[0]
    assert zero["commitments"]["F"] == bytes(64) and zero["commitments"]["T"] == bytes(64)


def test_many_columns(curve, tau, ptau_factory):
    """the reference accepts any number of columns (prover.js:34-45): k = 11 (one fused linear combination), k = 16
    (34 opening terms: the combinations go in chunks of 28, 32 round-1 commitments go in two result batches) and, for
    the grand product, k = 30"""
    n = 8
    for kind, k in (("gs", 11), ("gs", 16), ("gp", 30)):
        cols_f = [inputs.random_column(500 + i, n) for i in range(k)]
        cols_t = [inputs.rotate_right(c) for c in cols_f]
        got, want = _prove_pair(kind, curve, tau, ptau_factory, 3, cols_f, cols_t)
        assert pr.proof_bytes(got) == pr.proof_bytes(want), (kind, k)
        assert list(got["commitments"]) == list(want["commitments"])


def test_prover_replaces_evals_by_montgomery_form(curve, tau, ptau_factory):
    """prover.js:147-148: evalsFs[i].eval / evalsTs[i].eval are replaced by their Montgomery form"""
    from kzg_grandsums_study_b200.grandsum import mset_eq_kzg_grandsum_prover
    from kzg_grandsums_study_b200.polynomial import Evaluations
    n = 32
    cols = [inputs.random_column(70 + i, n) for i in range(2)]
    evf = [Evaluations(bn.fr_vec_to_std_bytes(c), curve) for c in cols]
    evt = [Evaluations(bn.fr_vec_to_std_bytes(inputs.rotate_right(c)), curve) for c in cols]
    mset_eq_kzg_grandsum_prover(ptau_factory(5), evf, evt)
    for c, ef, et in zip(cols, evf, evt):
        assert ef.tobytes() == bn.fr_vec_to_mont_bytes(c)
        assert et.tobytes() == bn.fr_vec_to_mont_bytes(inputs.rotate_right(c))
        assert ef.length() == n and ef.getEvaluation(3) == bn.fr_to_mont_bytes(c[3])


@pytest.mark.parametrize("kind", ["gs", "gp"])
def test_batched_verification_on_device(kind, curve, tau, ptau_factory):
    """m proofs, one device MSM per side of the pairing check over all their commitments, ONE pairing
    product.  24 honest proofs of mixed shapes are accepted; one bad proof anywhere makes the batch fail; the verdict
    agrees with the sequential drop-in verifier on every member."""
    from kzg_grandsums_study_b200.grandsum import mset_eq_kzg_grandsum_verifier_batch
    from kzg_grandsums_study_b200.grandproduct import mset_eq_kzg_grandproduct_verifier_batch
    batch_verify = mset_eq_kzg_grandsum_verifier_batch if kind == "gs" else mset_eq_kzg_grandproduct_verifier_batch
    nbits = 5
    proofs = []
    for j in range(24):
        k = 1 + j % 3
        got, _, _, _ = run_both(kind, curve, tau, ptau_factory, seed=900 + j, nbits=nbits, k=k, selected=(j % 4 == 3))
        proofs.append(got)
    path = run_both.last_path
    before = curve.launch_count()
    assert batch_verify(path, proofs, nbits) is True
    assert curve.launch_count() > before          # the group arithmetic ran on the device
    assert package_verify(kind, proofs[7], nbits) is True
    bad = [{"commitments": dict(p["commitments"]), "evaluations": dict(p["evaluations"])} for p in proofs]
    bad[17]["commitments"]["Q"] = bn.g1_to_bytes(bn.g1_mul_gen(5))
    assert batch_verify(path, bad, nbits) is False
    assert package_verify(kind, bad[17], nbits) is False
