"""CPU: pin the oracle against every known answer available for this path.

The reference's own tests hold no golden vectors (they assert only `isValid` on unseeded inputs), so the
pins are: public constants (Appendix B), the Keccak-256 KATs, the EIP-196 doubling vector, the derived
end-to-end vector of Appendix F, and self-consistency (closed-form commitments, verifier acceptance).
"""
import pytest

from oracle.py import bn254 as bn, inputs, keccak, poly, protocol as pr, ptau as pt


def test_field_constants():
    assert bn.Q == 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
    assert bn.R == 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
    assert bn.fr_to_mont_bytes(1).hex() == "fbffff4f1c3496ac29cd609f9576fc362e4679786fa36e662fdf079ac1770a0e"
    assert bn.fq_to_mont_bytes(1).hex() == "9d0d8fc58d435dd33d0bc7f528eb780a2c4679786fa36e662fdf079ac1770a0e"
    assert bn.fq_to_mont_bytes(2).hex() == "3a1b1e8b1b87baa67b168eeb51d6f114588cf2f0de46ddcc5ebe0f3483ef141c"


def test_roots_of_unity():
    w = bn.FR_W
    assert w[28] == 19103219067921713944291392827692070036145651957329286315305642004821462161904
    assert w[1] == bn.R - 1
    assert w[8] == 3478517300119284901893091970156912948790432420133812234316178878452092729974
    assert w[11] == 1120550406532664055539694724667294622065367841900378087843176726913374367458
    assert w[16] == 421743594562400382753388642386256516545992082196004333756405989743524594615
    assert w[20] == 17220337697351015657950521176323262483320249231368149235373741788599650842711
    for k in range(1, 28):
        assert w[k] == w[k + 1] * w[k + 1] % bn.R
        assert pow(w[k], 1 << k, bn.R) == 1 and pow(w[k], 1 << (k - 1), bn.R) != 1


def test_eip196_double():
    two_g = bn.g1_mul((1, 2), 2)
    assert two_g == (1368015179489954701390400359078579693043519447331113978918064868415326638035,
                     9918110051302171585080402603319702774565515993150576347155970296011118125764)
    assert bn.g1_is_on_curve(two_g)
    assert bn.g1_mul((1, 2), bn.R) is None


def test_keccak_kats():
    assert keccak.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    # Ethereum function selector of transfer(address,uint256)
    assert keccak.keccak256(b"transfer(address,uint256)").hex()[:8] == "a9059cbb"
    assert keccak.keccak256(b"a" * 200).hex() == keccak.keccak256(bytes([0x61]) * 200).hex()


def test_transcript_vectors():
    tr = pt.Keccak256Transcript()
    tr.add_pol_commitment(bn.g1_to_bytes((1, 2)))
    c1 = tr.get_challenge()
    assert bn.fr_from_mont_bytes(c1) == 17856212038068422348937662473302114032147350344021172871924595963388108456668
    tr.add_field_element(c1)
    assert bn.fr_from_mont_bytes(tr.get_challenge()) == \
        17784148524378679842293619980250942645783968316118172692717548276302528293647


def test_ntt_matches_definition():
    n = 16
    a = inputs.random_column(7, n)
    w = bn.FR_W[4]
    direct = [sum(a[j] * pow(w, j * k, bn.R) for j in range(n)) % bn.R for k in range(n)]
    assert poly.ntt(a) == direct
    assert poly.ntt(poly.ntt(a), inverse=True) == a


def test_batch_inverse_zero_maps_to_zero():
    v = [3, 0, 5, 0, bn.R - 1]
    inv = poly.batch_inverse(v)
    assert inv[1] == 0 and inv[3] == 0
    assert inv[0] * 3 % bn.R == 1 and inv[2] * 5 % bn.R == 1 and inv[4] == bn.R - 1


APPENDIX_F = {
    "F": (16835689094286762861045331424151979955283469995984224058260569358638687710068,
          1249451171975484176963229482762868423053517465179395881278980125831610377232),
    "T": (8229792471434702284516526369183808518504326039735693032834603020621756626773,
          94511063571212549527383130826953547854277962519665769126307322075172257548),
    "S": (16371981562895258283710529696859164115319383086406469448519223499104461797083,
          19129810415420991886844836059432180396867677759650705634434919156069294851673),
    "Q": (3846170189578177247615711312403726107231500294661943797203727171184267251888,
          16139925616325643864016290516984783167703223294533280720602985692920608701720),
    "Wxi": (11247552889987666679833405186666087740273206810747422399245363970622220056286,
            1219066491544469657863035456884313779234669039295406395184941652427545092387),
    "Wxiw": (10319162343374394764343463619286087936657639163999839545968593063611593848571,
             21696086493473458144269509860650732329214412569332322931926248127506125646569),
    "fxi": 8381979403735699037382406589247205630026910160665486970113847466528742283096,
    "txi": 9571836977897385062884946798335279722803481091970749539776705107405109633046,
    "sxiw": 3812625113219582593042459501361184331188234103107494215927524191815485429995,
    "gamma": 3276071517284392717626136523489035476000562272833661627600899621785393127826,
    "alpha": 19415459295846430477411834804415957918328311167825280710546251569079440937898,
    "xi": 11863047225567797371731950087428821411583900460366071224451582838280968918311,
    "v": 1823732689693580648281180238859153965614973555913571691917486416569593864988,
    "u": 14723220832594256613952603593450678686571775597920959564054087359314371649997,
}


def test_appendix_f_vector(tmp_path):
    """plain grand-sum, n = 4, tau = 0x1234567 -- with a real ptau file and a real MSM."""
    tau = 0x1234567
    path = str(tmp_path / "f.ptau")
    pt.write_ptau(path, 2, tau)
    srs = pr.Srs(path, 8)
    F = bn.fr_vec_to_std_bytes([1, 2, 3, 4])
    T = bn.fr_vec_to_std_bytes([4, 1, 2, 3])
    trace = {}
    proof = pr.grandsum_prover(srs, [F], [T], trace=trace)
    for k in ("F", "T", "S", "Q", "Wxi", "Wxiw"):
        assert bn.g1_from_bytes(proof["commitments"][k]) == APPENDIX_F[k], k
    for k in ("fxi", "txi", "sxiw"):
        assert bn.fr_from_mont_bytes(proof["evaluations"][k]) == APPENDIX_F[k], k
    for k in ("gamma", "alpha", "xi", "v"):
        assert trace["challenges"][k] == APPENDIX_F[k], k
    assert list(proof["commitments"]) == ["F", "T", "S", "Q", "Wxi", "Wxiw"]
    assert list(proof["evaluations"]) == ["fxi", "txi", "sxiw"]
    assert proof["commitments"]["F"].hex() == (
        "cd055c2b428e58a495972339985546dba726573ea29b6270ade6ec357fbaa204"
        "a504b976f13a4e456864b995f8e82c337678c53710968db0930eb9f60323dc25")
    assert proof["evaluations"]["fxi"].hex() == "6228f33f6983e1f743118a275b003d1617857745952bd101192d11346285be1c"
    ch = {}
    assert pr.grandsum_verifier(proof, 2, tau=tau, out_challenges=ch)
    assert ch["u"] == APPENDIX_F["u"]
    # the closed form agrees with the MSM
    proof2 = pr.grandsum_prover(pr.TrapdoorSrs(tau, 2), [F], [T])
    assert pr.proof_bytes(proof2) == pr.proof_bytes(proof)


@pytest.mark.parametrize("kind", ["gs", "gp"])
@pytest.mark.parametrize("k,selected", [(1, False), (3, False), (1, True), (2, True)])
def test_oracle_self_consistency(kind, k, selected):
    """honest proofs verify (the property the reference's tests assert, test/mset_eq_kzg_grandsum.test.js:24-104);
    a tampered proof does not"""
    nbits = 3
    n = 1 << nbits
    tau = inputs.tau_from_seed(99)
    cols_f = [inputs.random_column(10 + i, n) for i in range(k)]
    perm = inputs.permutation(5, n)
    sel_f = sel_t = None
    if selected:
        # reference test construction: T = F rotated right, selF[n-1] = 0, selT[0] = 0
        cols_t = [inputs.rotate_right(c) for c in cols_f]
        one = bn.fr_to_mont_bytes(1)
        zero = bytes(32)
        sel_f = one * (n - 1) + zero
        sel_t = zero + one * (n - 1)
    else:
        cols_t = [[c[perm[i]] for i in range(n)] for c in cols_f]
    prover = pr.grandsum_prover if kind == "gs" else pr.grandproduct_prover
    verifier = pr.grandsum_verifier if kind == "gs" else pr.grandproduct_verifier
    proof = prover(pr.TrapdoorSrs(tau, nbits), [bn.fr_vec_to_std_bytes(c) for c in cols_f],
                   [bn.fr_vec_to_std_bytes(c) for c in cols_t], sel_f, sel_t)
    assert verifier(proof, nbits, tau=tau)
    key = next(iter(proof["evaluations"]))
    bad = bn.fr_to_mont_bytes((bn.fr_from_mont_bytes(proof["evaluations"][key]) + 1) % bn.R)
    proof["evaluations"][key] = bad
    assert not verifier(proof, nbits, tau=tau)


def test_unequal_multisets_rejected():
    n = 8
    f = inputs.random_column(1, n)
    t = list(f)
    t[3] = (t[3] + 1) % bn.R
    with pytest.raises(Exception, match="not well calculated"):
        pr.grandsum_prover(pr.TrapdoorSrs(5, 3), [bn.fr_vec_to_std_bytes(f)], [bn.fr_vec_to_std_bytes(t)])
    with pytest.raises(Exception, match="not well calculated"):
        pr.grandproduct_prover(pr.TrapdoorSrs(5, 3), [bn.fr_vec_to_std_bytes(f)], [bn.fr_vec_to_std_bytes(t)])


def test_oracle_reproduces_golden_fixtures():
    """the committed fixtures (tests/golden/*.json) are what the oracle produces today"""
    import json
    import os
    import sys
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, golden)
    import make_golden
    tau = inputs.tau_from_seed(make_golden.TAU_SEED)
    names = sorted(f for f in os.listdir(golden) if f.endswith(".json"))
    assert len(names) >= 8
    for name in names:
        g = json.load(open(os.path.join(golden, name)))
        cf, ct, sf, st = make_golden.columns(g["seed"], g["nbits"], g["k"], g["selected"], g["rotate"])
        prover = pr.grandsum_prover if g["kind"] == "gs" else pr.grandproduct_prover
        proof = prover(pr.TrapdoorSrs(tau, g["ptau_power"]), [bn.fr_vec_to_std_bytes(c) for c in cf],
                       [bn.fr_vec_to_std_bytes(c) for c in ct], sf, st)
        assert pr.proof_bytes(proof).hex() == g["proof_bytes"], name


def test_reference_polynomial_kats():
    """the small-integer known answers of the reference's test/polynomial.test.js for methods on the prover path:
    evaluate (:116-124) and multiply (:207-220)"""
    P = poly.Polynomial
    assert P([0, 1, 2, 3]).evaluate(2) == 34
    prod = P([2, 0, bn.R - 3, 2]).multiply(P([0, 3, 1]))
    assert prod.coef[:6] == [0, 6, 2, bn.R - 9, 3, 2] and not any(prod.coef[6:])
    # divByXSubValue inverts byXSubValue's KAT (:255-262): (7x^2 - 3x + 4)(x - 6) = 7x^3 - 45x^2 + 22x - 24
    q = P([bn.R - 24, 22, bn.R - 45, 7]).div_by_x_sub_value(6)
    assert q.coef == [4, bn.R - 3, 7, 0]
