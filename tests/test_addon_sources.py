"""CPU: the N-API addon and the JS drop-in layer are shipped as source (north_star: "Node host code calls a thin C-ABI N-API
addon").  This image has no Node, so they are checked statically:
  * addon/kzgb200_napi.cc is exactly what tools/gen_napi.py generates from include/kzgb200.h, binds EVERY declared symbol, and
    type-checks with g++ against a stub node_api.h that carries Node's real signatures;
  * every `<addon>.kzg_*(...)` call in js/ names a generated binding and passes the number of arguments it takes;
  * js/ mirrors the reference's module layout and entry-point signatures, brackets balance, and the proof keys are written
    in the reference's insertion order."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _gen():
    import gen_napi
    return gen_napi


def test_addon_is_generated_from_the_header():
    g = _gen()
    text = open(os.path.join(ROOT, "addon", "kzgb200_napi.cc")).read()
    assert text == g.generate(), "addon/kzgb200_napi.cc is stale: run python tools/gen_napi.py"
    names = [n for _, n, _ in g.declarations()]
    assert len(names) >= 90
    for n in names:
        assert '{"%s", js_%s},' % (n, n) in text, n


def test_addon_type_checks_against_stub_node_api():
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "addon", "stub"), "-I",
           os.path.join(ROOT, "include"), os.path.join(ROOT, "addon", "kzgb200_napi.cc")]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout[-3000:]


def _js_files():
    out = []
    for base, _, files in os.walk(os.path.join(ROOT, "js")):
        out += [os.path.join(base, f) for f in files if f.endswith(".js")]
    for base, _, files in os.walk(os.path.join(ROOT, "bench", "ref_node")):
        out += [os.path.join(base, f) for f in files if f.endswith(".js")]
    return sorted(out)


def _strip(text):
    """remove comments, strings, template literals and regex literals well enough for bracket / call analysis"""
    out, i, n = [], 0, len(text)
    while i < n:
        c = text[i]
        if text.startswith("//", i):
            i = text.find("\n", i)
            i = n if i < 0 else i
        elif text.startswith("/*", i):
            i = text.find("*/", i) + 2
        elif c in "\"'`":
            q = c
            i += 1
            depth = 0
            while i < n and (text[i] != q or depth):
                if text[i] == "\\":
                    i += 1
                elif q == "`" and text.startswith("${", i):
                    depth += 1
                    i += 1
                elif q == "`" and depth and text[i] == "}":
                    depth -= 1
                i += 1
            i += 1
            out.append('""')
        elif c == "/" and re.search(r"[=(,:!&|?;{}\[]\s*$", "".join(out[-40:]) or "("):
            i += 1                                   # regex literal
            while i < n and text[i] != "/":
                i += 2 if text[i] == "\\" else 1
            i += 1
            while i < n and text[i].isalpha():
                i += 1
            out.append('""')
        else:
            out.append(c)
            i += 1
    return "".join(out)


@pytest.mark.parametrize("path", _js_files(), ids=lambda p: os.path.relpath(p, ROOT))
def test_js_brackets_balance(path):
    text = _strip(open(path).read())
    stack = []
    pairs = {")": "(", "]": "[", "}": "{"}
    for ch in text:
        if ch in "([{":
            stack.append(ch)
        elif ch in ")]}":
            assert stack and stack.pop() == pairs[ch], "unbalanced %r in %s" % (ch, path)
    assert not stack


def _count_args(text, start):
    """text[start] == '(' -> number of top-level arguments"""
    depth, args, seen = 0, 0, False
    i = start
    while i < len(text):
        ch = text[i]
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
            if depth == 0:
                return args + (1 if seen else 0)
        elif ch == "," and depth == 1:
            args += 1
        elif not ch.isspace() and depth >= 1:
            seen = True
        i += 1
    raise AssertionError("unterminated call")


def test_js_calls_match_the_generated_bindings():
    g = _gen()
    arity = {}
    for ret, name, params in g.declarations():
        arity[name] = sum(1 for p in params if g.classify(name, *p)[0] in ("handle", "string", "bytes", "address", "scalar", "array"))
    calls = 0
    for path in _js_files():
        text = _strip(open(path).read())
        for m in re.finditer(r"\b(?:a|addon|curve\.addon|this\.curve\.addon|this\._curve\.addon|this\.addon)\.(kzg_[a-z0-9_]+)\s*\(", text):
            name = m.group(1)
            assert name in arity, "%s calls unknown binding %s" % (os.path.relpath(path, ROOT), name)
            got = _count_args(text, m.end() - 1)
            assert got == arity[name], "%s: %s called with %d argument(s), the binding takes %d" % (
                os.path.relpath(path, ROOT), name, got, arity[name])
            calls += 1
        for m in re.finditer(r"addon\[(fn)\]|\.addon\[fn\]", text):
            calls += 0
    assert calls >= 30


def test_js_layout_and_signatures_mirror_the_reference():
    want = ["index.js", "src/curve.js", "src/Keccak256Transcript.js", "src/ptau_utils.js", "src/polynomial/polynomial.js",
            "src/polynomial/evaluations.js", "src/polynomial/polynomial_utils.js", "src/grandsum/mset_eq_kzg_prover.js",
            "src/grandsum/mset_eq_kzg_verifier.js", "src/grandsum/grandsum.js", "src/grandproduct/mset_eq_kzg_prover.js",
            "src/grandproduct/mset_eq_kzg_verifier.js", "src/grandproduct/grandproduct.js"]
    for rel in want:
        assert os.path.exists(os.path.join(ROOT, "js", rel)), rel
    src = lambda rel: open(os.path.join(ROOT, "js", rel)).read()
    # entry points: names and argument lists of prover.js:12 / verifier.js:9 / grandsum.js:6 / grandproduct.js:6
    assert "async function mset_eq_kzg_grandsum_prover(pTauFilename, evalsFs, evalsTs, evalsSelF = null, evalsSelT = null" in src("src/grandsum/mset_eq_kzg_prover.js")
    assert "async function mset_eq_kzg_grandproduct_prover(pTauFilename, evalsFs, evalsTs, evalsSelF = null, evalsSelT = null" in src("src/grandproduct/mset_eq_kzg_prover.js")
    assert "async function mset_eq_kzg_grandsum_verifier(pTauFilename, proof, nBits" in src("src/grandsum/mset_eq_kzg_verifier.js")
    assert "async function mset_eq_kzg_grandproduct_verifier(pTauFilename, proof, nBits" in src("src/grandproduct/mset_eq_kzg_verifier.js")
    assert "ComputeSGrandSumPolynomial(evalsF, evalsT, evalsSelF, evalsSelT, challenge, curve)" in src("src/grandsum/grandsum.js")
    assert "ComputeZGrandProductPolynomial(evalsF, evalsT, evalsSelF, evalsSelT, isSelected, challenge, curve)" in src("src/grandproduct/grandproduct.js")
    assert "async function readPTauHeader(fd, sections" in src("src/ptau_utils.js")
    poly = src("src/polynomial/polynomial.js")
    for method in ("static async fromEvaluations(", "static fromCoefficientsArray(", "static fromPolynomial(", "static zero(",
                   "static async Lagrange1(", "clone()", "isEqual(", "getCoef(", "setCoef(", "length()", "degree()", "evaluate(",
                   "add(polynomial, blindingValue)", "sub(polynomial, blindingValue)", "async multiply(", "async shiftOmega()",
                   "mulScalar(", "addScalar(", "subScalar(", "divByXSubValue(", "divZh(", "async multiExponentiation(PTau, name)"):
        assert method in poly, method
    assert "this.coef = coefficients" in poly and "this.eval = evaluations" in src("src/polynomial/evaluations.js")
    # proof keys in the reference's insertion order (prover.js:161-162,173-174,229,284,301-316,409-410)
    drv = _strip(src("src/prover_common.js"))
    order = [drv.index(tok) for tok in ("Cm[fName(i)] =", "Cm.selF =", "Cm[acc] =", "Cm.Q =", "Cm.Wxi =", "Cm.Wxiw =")]
    assert order == sorted(order)
    # the reference's error strings (prover.js:30-81)
    for msg in ("The lengths of the two vector multisets must be the same.", "The number of multisets must be greater than 0.",
                "The multiset buffers must all have the same length.", "The selection buffers must have the same length.",
                "Polynomial length must be a power of two.",
                "The Powers of Tau file is not sufficiently large to commit the polynomials."):
        assert msg in src("src/prover_common.js"), msg
