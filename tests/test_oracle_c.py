"""CPU: the oracle's C restatement (oracle/c, the CPU baseline of bench.py) against the Python oracle,
the SURVEY.md Appendix F vector and the committed golden fixtures."""
import json
import os
import sys

import pytest

from oracle.c import binding as oc
from oracle.py import bn254 as bn, inputs, keccak, poly as opoly, protocol as pr

R = bn.R
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_field_and_bulk_ops():
    v = inputs.random_column(1, 300) + [0, 1, R - 1]
    std = bn.fr_vec_to_std_bytes(v)
    assert oc.to_mont(std) == bn.fr_vec_to_mont_bytes(v)
    assert oc.from_mont(bn.fr_vec_to_mont_bytes(v)) == std
    w = list(v)
    w[5] = 0
    assert bn.fr_vec_from_mont_bytes(oc.batch_inverse(bn.fr_vec_to_mont_bytes(w))) == opoly.batch_inverse(w)
    for data in (b"", b"abc", bytes(range(200))):
        assert oc.keccak256(data) == keccak.keccak256(data)


@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 10, 13])
def test_ntt(log_n):
    a = inputs.random_column(3 + log_n, 1 << log_n)
    m = bn.fr_vec_to_mont_bytes(a)
    assert bn.fr_vec_from_mont_bytes(oc.ntt(m)) == opoly.ntt(a)
    assert bn.fr_vec_from_mont_bytes(oc.ntt(m, inverse=True)) == opoly.ntt(a, inverse=True)


def test_srs_and_msm():
    tau = inputs.tau_from_seed(1001)
    n = 600
    srs = oc.srs_generate(tau, n)
    t = 1
    for i in range(20):
        assert srs[64 * i:64 * i + 64] == bn.g1_to_bytes(bn.g1_mul_gen(t))
        t = t * tau % R
    assert oc.srs_generate(tau, 8, first=500) == srs[64 * 500:64 * 508]
    for m in (1, 2, 17, 600):
        s = inputs.random_column(40 + m, m)
        if m > 2:
            s[0], s[1] = 0, R - 1
        expect = sum(x * pow(tau, i, R) for i, x in enumerate(s)) % R
        for threads in (1, 4):
            assert oc.msm(srs[:64 * m], bn.fr_vec_to_std_bytes(s), threads) == bn.g1_to_bytes(bn.g1_mul_gen(expect))
    assert oc.msm(b"", b"") == bytes(64)
    G = bn.g1_to_bytes((1, 2))
    assert oc.msm(G + G, bn.fr_vec_to_std_bytes([3, 4])) == bn.g1_to_bytes(bn.g1_mul_gen(7))
    assert oc.msm(G + bn.g1_to_bytes((1, bn.Q - 2)), bn.fr_vec_to_std_bytes([5, 5])) == bytes(64)


def test_appendix_f():
    tau = 0x1234567
    srs = oc.srs_generate(tau, 8)
    proof, ch = oc.prove("gs", srs, [bn.fr_vec_to_std_bytes([1, 2, 3, 4])], [bn.fr_vec_to_std_bytes([4, 1, 2, 3])])
    assert proof[:64].hex() == ("cd055c2b428e58a495972339985546dba726573ea29b6270ade6ec357fbaa204"
                                "a504b976f13a4e456864b995f8e82c337678c53710968db0930eb9f60323dc25")
    assert proof[64 * 6:64 * 6 + 32].hex() == "6228f33f6983e1f743118a275b003d1617857745952bd101192d11346285be1c"
    assert bn.fr_from_mont_bytes(ch["v"]) == 1823732689693580648281180238859153965614973555913571691917486416569593864988
    want = pr.grandsum_prover(pr.TrapdoorSrs(tau, 2), [bn.fr_vec_to_std_bytes([1, 2, 3, 4])], [bn.fr_vec_to_std_bytes([4, 1, 2, 3])])
    assert proof == pr.proof_bytes(want)


def test_golden_fixtures():
    sys.path.insert(0, GOLDEN)
    import make_golden
    tau = inputs.tau_from_seed(make_golden.TAU_SEED)
    srs_cache = {}
    for name in sorted(f for f in os.listdir(GOLDEN) if f.endswith(".json")):
        g = json.load(open(os.path.join(GOLDEN, name)))
        n = 1 << g["nbits"]
        if n not in srs_cache:
            srs_cache[n] = oc.srs_generate(tau, 2 * n)
        cf, ct, sf, st = make_golden.columns(g["seed"], g["nbits"], g["k"], g["selected"], g["rotate"])
        proof, ch = oc.prove(g["kind"], srs_cache[n], [bn.fr_vec_to_std_bytes(c) for c in cf],
                             [bn.fr_vec_to_std_bytes(c) for c in ct], sf, st)
        assert proof.hex() == g["proof_bytes"], name
        for nm, val in g["challenges"].items():
            assert bn.fr_from_mont_bytes(ch[nm]) == int(val, 16), (name, nm)


def test_errors():
    tau = 5
    srs = oc.srs_generate(tau, 16)
    f = inputs.random_column(1, 8)
    t = list(f)
    t[3] = (t[3] + 1) % R
    for kind in ("gs", "gp"):
        with pytest.raises(ValueError, match="not well calculated"):
            oc.prove(kind, srs, [bn.fr_vec_to_std_bytes(f)], [bn.fr_vec_to_std_bytes(t)])
    two, one = bn.fr_to_mont_bytes(2), bn.fr_to_mont_bytes(1)
    with pytest.raises(ValueError, match="not divisible"):
        oc.prove("gs", srs, [bn.fr_vec_to_std_bytes(f)], [bn.fr_vec_to_std_bytes(f)], two + one * 7, two + one * 7)
