"""CPU: the drop-in verifiers (host code with a real optimal-ate pairing) accept the oracle's honest proofs and reject
tampered ones, in agreement with the oracle's trapdoor check -- the property the reference's tests assert
(test/mset_eq_kzg_grandsum.test.js:24-104: prove, verify, assert.ok(isValid))."""
import pytest

from oracle.py import bn254 as bn, inputs, protocol as pr, ptau as opt


def test_pairing_bilinear(lib_path):
    from kzg_grandsums_study_b200 import host_bn254 as hb
    e1 = hb.pairing(hb.G2_GEN, hb.G1_GEN)
    assert e1 != hb._F12_ONE and hb.f12_pow(e1, hb.R) == hb._F12_ONE
    assert hb.pairing(hb.G2_GEN, hb.g1_mul(hb.G1_GEN, 3)) == hb.f12_pow(e1, 3)
    assert hb.pairing(bn.g2_mul(bn.G2_GEN, 5), hb.G1_GEN) == hb.f12_pow(e1, 5)
    tau, a = 123456789, 987654321
    A = hb.g1_neg(hb.g1_mul(hb.G1_GEN, a))
    B = hb.g1_mul(hb.G1_GEN, a * tau % hb.R)
    assert hb.pairing_eq(A, bn.g2_mul(bn.G2_GEN, tau), B, hb.G2_GEN)
    assert not hb.pairing_eq(A, bn.g2_mul(bn.G2_GEN, tau + 1), B, hb.G2_GEN)
    # host G1 arithmetic against the oracle
    for k in (1, 2, 7, bn.R - 1, 1 << 200):
        assert hb.g1_mul(hb.G1_GEN, k) == bn.g1_mul((1, 2), k)


@pytest.mark.parametrize("kind", ["gs", "gp"])
@pytest.mark.parametrize("k,selected", [(1, False), (3, False), (1, True), (2, True)])
def test_verifiers_accept_and_reject(kind, k, selected, tmp_path, lib_path):
    from kzg_grandsums_study_b200.grandsum import mset_eq_kzg_grandsum_verifier
    from kzg_grandsums_study_b200.grandproduct import mset_eq_kzg_grandproduct_verifier
    nbits = 3
    n = 1 << nbits
    tau = inputs.tau_from_seed(4242)
    path = str(tmp_path / "v.ptau")
    opt.write_ptau(path, nbits, tau)
    cols_f = [inputs.random_column(10 + i, n) for i in range(k)]
    cols_t = [inputs.rotate_right(c) for c in cols_f]
    sel_f = sel_t = None
    if selected:
        one, zero = bn.fr_to_mont_bytes(1), bytes(32)
        sel_f, sel_t = one * (n - 1) + zero, zero + one * (n - 1)
    prover = pr.grandsum_prover if kind == "gs" else pr.grandproduct_prover
    oracle_verifier = pr.grandsum_verifier if kind == "gs" else pr.grandproduct_verifier
    verifier = mset_eq_kzg_grandsum_verifier if kind == "gs" else mset_eq_kzg_grandproduct_verifier
    proof = prover(pr.Srs(path, 2 * n), [bn.fr_vec_to_std_bytes(c) for c in cols_f],
                   [bn.fr_vec_to_std_bytes(c) for c in cols_t], sel_f, sel_t)
    assert verifier(path, proof, nbits) is True
    assert oracle_verifier(proof, nbits, tau=tau)
    # tamper with an evaluation, then with a commitment
    key = next(iter(proof["evaluations"]))
    good = proof["evaluations"][key]
    proof["evaluations"][key] = bn.fr_to_mont_bytes((bn.fr_from_mont_bytes(good) + 1) % bn.R)
    assert verifier(path, proof, nbits) is False
    proof["evaluations"][key] = good
    good_c = proof["commitments"]["Q"]
    proof["commitments"]["Q"] = bn.g1_to_bytes(bn.g1_mul_gen(5))
    assert verifier(path, proof, nbits) is False
    # not a curve point / not a field element: rejected without raising (verifier.js:50,61)
    proof["commitments"]["Q"] = bn.fq_to_mont_bytes(1) + bn.fq_to_mont_bytes(1)
    assert verifier(path, proof, nbits) is False
    proof["commitments"]["Q"] = good_c
    proof["evaluations"][key] = b"\xff" * 32
    assert verifier(path, proof, nbits) is False
    proof["evaluations"][key] = good
    assert verifier(path, proof, nbits) is True
    # wrong domain size
    assert verifier(path, proof, nbits + 1) is False
    # a non-canonical encoding of a good commitment (x + q when it fits 256 bits) and a selected proof that lost its
    # selector evaluations: both refused without raising
    raw = proof["commitments"]["Q"]
    x = int.from_bytes(raw[:32], "little")
    if x + bn.Q < 1 << 256:
        proof["commitments"]["Q"] = (x + bn.Q).to_bytes(32, "little") + raw[32:]
        assert verifier(path, proof, nbits) is False
        proof["commitments"]["Q"] = raw
    if selected:
        lost = proof["evaluations"].pop("selFxi")
        assert verifier(path, proof, nbits) is False
        proof["evaluations"]["selFxi"] = lost


def test_verifier_refuses_non_canonical_and_incomplete_proofs():
    """a coordinate encoded as x + q names the same point but hashes differently -- refused; a selected proof
    without selFxi / selTxi is refused, not a KeyError"""
    from kzg_grandsums_study_b200 import host_bn254 as hb
    g = hb.g1_to_bytes((1, 2))
    assert hb.g1_bytes_canonical(g) and hb.g1_bytes_canonical(bytes(64))
    x = int.from_bytes(g[:32], "little")
    if x + hb.Q < 1 << 256:
        shifted = (x + hb.Q).to_bytes(32, "little") + g[32:]
        assert hb.g1_from_bytes(shifted) == hb.g1_from_bytes(g)
        assert not hb.g1_bytes_canonical(shifted)
    assert not hb.g1_bytes_canonical(b"\xff" * 64)


def _host_msm(bases, scalars):
    """a host multi-scalar multiplication for the CPU test of the batching logic (the product uses the device MSM)"""
    from kzg_grandsums_study_b200 import host_bn254 as hb
    acc = None
    for i in range(len(scalars) // 32):
        P = hb.g1_from_bytes(bases[64 * i:64 * i + 64])
        k = int.from_bytes(scalars[32 * i:32 * i + 32], "little")
        acc = hb.g1_add(acc, hb.g1_mul(P, k))
    return hb.g1_to_bytes(acc)


@pytest.mark.parametrize("kind", ["gs", "gp"])
def test_batch_verifier_logic(kind, tmp_path, lib_path):
    """the linear forms of verify_batch reproduce the sequential verifier: a batch of honest proofs of different shapes
    (plain, vector, selected) is accepted, the same batch with ONE tampered proof anywhere is rejected, and the single
    linear forms give the same pairing inputs as steps 6-9 of verify()"""
    from kzg_grandsums_study_b200 import _verifier_common as vc, host_bn254 as hb
    nbits = 3
    n = 1 << nbits
    tau = inputs.tau_from_seed(777)
    path = str(tmp_path / "b.ptau")
    opt.write_ptau(path, nbits, tau)
    prover = pr.grandsum_prover if kind == "gs" else pr.grandproduct_prover
    one, zero = bn.fr_to_mont_bytes(1), bytes(32)
    proofs = []
    for k, selected in ((1, False), (3, False), (2, True), (1, True)):
        cols_f = [inputs.random_column(300 + 10 * k + i, n) for i in range(k)]
        cols_t = [inputs.rotate_right(c) for c in cols_f]
        sel = (one * (n - 1) + zero, zero + one * (n - 1)) if selected else (None, None)
        proofs.append(prover(pr.Srs(path, 2 * n), [bn.fr_vec_to_std_bytes(c) for c in cols_f],
                             [bn.fr_vec_to_std_bytes(c) for c in cols_t], sel[0], sel[1]))
    # single proofs through the linear forms == the sequential verifier's verdict
    for proof in proofs:
        assert vc.verify(kind, path, proof, nbits) is True
        assert vc.verify_batch(kind, path, [proof], nbits, msm=_host_msm) is True
    assert vc.verify_batch(kind, path, proofs, nbits, msm=_host_msm) is True
    assert vc.verify_batch(kind, path, [], nbits, msm=_host_msm) is True
    for bad_at in range(len(proofs)):
        batch = [{"commitments": dict(p["commitments"]), "evaluations": dict(p["evaluations"])} for p in proofs]
        key = list(batch[bad_at]["evaluations"])[-1]
        good = batch[bad_at]["evaluations"][key]
        batch[bad_at]["evaluations"][key] = bn.fr_to_mont_bytes((bn.fr_from_mont_bytes(good) + 1) % bn.R)
        assert vc.verify_batch(kind, path, batch, nbits, msm=_host_msm) is False, bad_at
        batch[bad_at]["evaluations"][key] = good
        batch[bad_at]["commitments"]["Wxi"] = bn.g1_to_bytes(bn.g1_mul_gen(11 + bad_at))
        assert vc.verify_batch(kind, path, batch, nbits, msm=_host_msm) is False, bad_at
    # malformed member: refused without raising
    batch = [proofs[0], {"commitments": {}, "evaluations": {}}]
    assert vc.verify_batch(kind, path, batch, nbits, msm=_host_msm) is False
