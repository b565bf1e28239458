"""Generator + word-level emulator of the dedicated Montgomery squaring in csrc/field.cuh (fp_sqr, device fast path).

    python tools/gen_fp_sqr.py check       emulate the instruction list on random / edge operands against a*a*R^-1 mod p
    python tools/gen_fp_sqr.py emit        print the CUDA body (the block between the GENERATED markers of field.cuh)
    python tools/gen_fp_sqr.py update      rewrite that block in place

The squaring is 100 wide multiply-adds (mad.lo.cc + madc.hi.cc pairs = one IMAD.WIDE.U32 each in SASS) where the general
product spends 128:  28 off-diagonal products a_i a_j (i < j) in two accumulators -- E takes the products whose limb position
i + j is even, O the odd ones, so that every row is ONE carry chain over consecutive word pairs --, S = E + (O << 32),
T = 2 S + sum a_i^2 2^(64 i) (funnel shifts + one 16-word chain, 8 wide MACs), then a Montgomery reduction of the 512-bit T
with the even/odd row scheme of fp_mul (64 wide MACs), + T_high, one conditional subtraction.

Every instruction is a tuple in PTX order; a `block` is one asm statement (the carry flag never crosses a statement).  The
SAME list is emulated here and printed as inline PTX, so that the carry bookkeeping is checked on the CPU (there is no GPU
in the build container); kzg_selftest() then checks the compiled code against fp_mul_portable on the device.
"""
import random
import sys
import os

Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R_ = 21888242871839275222246405745257275088548364400416034343698204186575808495617
M32 = 0xFFFFFFFF


class Block:
    def __init__(self):
        self.ins = []   # (op, dst, srcs...)  srcs are variable names, ints (immediates) or ('mod', k) / ('inv',)

    def add(self, *t):
        self.ins.append(t)


def append_redc(blocks, T):
    """Montgomery reduction of the 16-word T (< 2 p^2 + ...): appends the blocks that leave t0..t7 = T / 2^256 mod p, < 2 p"""
    # ---- Montgomery reduction of T: U = (T_low + sum m_k p 2^(32 k)) / 2^256 <= p, result = U + T_high < 2 p
    # pair (lo: positions 0..7, hi: positions 1..8); after each step the total is divisible by 2^32 and the roles swap
    x = ["x%d" % k for k in range(8)]
    y = ["y%d" % k for k in range(8)]
    mod = [("mod", k) for k in range(8)]
    b = Block()
    for k in range(8):
        b.add("mov", x[k], T[k])
    b.add("mul.lo", "m", x[0], ("inv",))
    blocks.append(b)
    b = Block()
    for n, k in enumerate((1, 3, 5, 7)):   # y = odd limbs of p times m (plain products)
        b.add("mul.lo", y[2 * n], mod[k], "m")
        b.add("mul.hi", y[2 * n + 1], mod[k], "m")
    blocks.append(b)

    def cmad4_top(lo, m, top):
        b = Block()
        for n, k in enumerate((0, 2, 4, 6)):
            b.add(("mad.lo.cc" if n == 0 else "madc.lo.cc"), lo[2 * n], mod[k], m, lo[2 * n])
            b.add("madc.hi.cc", lo[2 * n + 1], mod[k], m, lo[2 * n + 1])
        b.add("addc", top, top, 0)
        blocks.append(b)

    cmad4_top(x, "m", y[7])
    for i in range(1, 8):
        lo, hi = (y, x) if i & 1 else (x, y)
        # the old lo (now `hi`): word 0 is zero, word 1 belongs to the new position 0, words 2..7 become the new hi 0..5
        b = Block()
        b.add("addw", "w", lo[0], hi[1])      # wrapping: only the low word decides m
        b.add("mul.lo", "m", "w", ("inv",))
        blocks.append(b)
        b = Block()
        b.add("add.cc", lo[0], lo[0], hi[1])
        src = [hi[2], hi[3], hi[4], hi[5], hi[6], hi[7], 0, 0]
        for n, k in enumerate((1, 3, 5, 7)):
            b.add("madc.lo.cc", hi[2 * n], mod[k], "m", src[2 * n])
            b.add("madc.hi" + (".cc" if n < 3 else ""), hi[2 * n + 1], mod[k], "m", src[2 * n + 1])
        blocks.append(b)
        cmad4_top(lo, "m", hi[7])
    # after step 7 (odd): lo = y, hi = x; result word j = hi[j] + lo[j + 1] (lo[0] == 0), then + T_high
    lo, hi = y, x
    b = Block()
    for j in range(8):
        op = "add.cc" if j == 0 else ("addc.cc" if j < 7 else "addc")
        b.add(op, "t%d" % j, hi[j], lo[j + 1] if j < 7 else 0)
    blocks.append(b)
    b = Block()
    for j in range(8):
        op = "add.cc" if j == 0 else ("addc.cc" if j < 7 else "addc")
        b.add(op, "t%d" % j, "t%d" % j, T[8 + j])
    blocks.append(b)


def build():
    """-> list of Blocks; variables: a0..a7 (inputs), t0..t7 (result before the final subtraction)."""
    blocks = []
    a = ["a%d" % i for i in range(8)]
    E = ["e%d" % i for i in range(16)]
    O = ["o%d" % i for i in range(16)]   # o[k] = word position k + 1

    written = set()

    def chain(acc, base, i, js, carry_to):
        """acc[base + ...] += a_i * a_j for j in js (consecutive word pairs starting at acc[base]); carry -> acc[carry_to]"""
        b = Block()
        first = True
        for n, j in enumerate(js):
            lo, hi = acc[base + 2 * n], acc[base + 2 * n + 1]
            for word, part in ((lo, "lo"), (hi, "hi")):
                addend = word if word in written else 0
                last = (n == len(js) - 1 and part == "hi")
                cc_out = not (last and carry_to is None)
                if first:
                    if addend == 0:
                        b.add("mul." + part, word, a[j], a[i])
                        # a plain mul sets no carry: the next op must not read one
                        nocarry = True
                    else:
                        b.add("mad.%s%s" % (part, ".cc" if cc_out else ""), word, a[j], a[i], addend)
                        nocarry = False
                    first = False
                else:
                    if nocarry:
                        if addend == 0:
                            b.add("mul." + part, word, a[j], a[i])
                        else:
                            b.add("mad.%s%s" % (part, ".cc" if cc_out else ""), word, a[j], a[i], addend)
                            nocarry = False
                    else:
                        b.add("madc.%s%s" % (part, ".cc" if cc_out else ""), word, a[j], a[i], addend)
                written.add(word)
        if carry_to is not None:
            w = acc[carry_to]
            assert w not in written, w
            assert not nocarry
            b.add("addc", w, 0, 0)
            written.add(w)
        blocks.append(b)

    # ---- off-diagonal products, rows by multiplier a_i; tops never decrease, so a carry out always lands in a word that
    # holds nothing but carries so far
    for i in range(7):
        ev = [j for j in range(i + 1, 8) if (i + j) % 2 == 0]
        od = [j for j in range(i + 1, 8) if (i + j) % 2 == 1]
        if ev:
            base = i + ev[0]
            top = base + 2 * len(ev) - 1            # last word the chain writes
            fresh_top = E[top] not in written
            chain(E, base, i, ev, None if fresh_top else top + 1)
        if od:
            base = i + od[0] - 1
            top = base + 2 * len(od) - 1
            fresh_top = O[top] not in written
            chain(O, base, i, od, None if fresh_top else top + 1)
    # E holds positions 2..13, O positions 1..14 (o0..o13)
    assert all(("e%d" % k in written) == (2 <= k <= 13) for k in range(16)), sorted(written)
    assert all(("o%d" % k in written) == (k <= 13) for k in range(16)), sorted(written)

    # ---- S = E + (O << 32): s[k] = e[k] + o[k-1]; s0 = 0, s1 = o0, s14 = o13 + carry, s15 = 0 (the sum is < 2^480)
    b = Block()
    S = [None] * 16
    S[0] = 0
    S[1] = "o0"
    for k in range(2, 14):
        op = "add.cc" if k == 2 else "addc.cc"
        b.add(op, E[k], E[k], O[k - 1])
        S[k] = E[k]
    b.add("addc", "o13", "o13", 0)
    S[14] = "o13"
    blocks.append(b)
    # ---- D = 2 S (funnel shifts), d0 = 0
    b = Block()
    D = ["d%d" % k for k in range(16)]
    for k in range(15, 0, -1):   # top down: in place would be possible, separate names keep the emulator simple
        lo = S[k - 1]
        hi = S[k] if k < 15 else 0
        if hi == 0:
            b.add("shr31", D[k], lo)
        elif lo == 0:
            b.add("shl1", D[k], hi)
        else:
            b.add("shf.l", D[k], lo, hi)   # (hi << 1) | (lo >> 31)
    blocks.append(b)
    # ---- T = D + sum a_i^2 2^(64 i): one chain over 16 words
    b = Block()
    T = ["T%d" % k for k in range(16)]
    for i in range(8):
        if i == 0:
            b.add("mul.lo", T[0], a[0], a[0])
            b.add("mad.hi.cc", T[1], a[0], a[0], D[1])
        else:
            b.add("madc.lo.cc", T[2 * i], a[i], a[i], D[2 * i])
            b.add("madc.hi" + (".cc" if i < 7 else ""), T[2 * i + 1], a[i], a[i], D[2 * i + 1])
    blocks.append(b)

    append_redc(blocks, T)
    return blocks


# ------------------------------------------------------------------------------------------------------------------
def emulate(blocks, aval, p):
    inv = (-pow(p, -1, 1 << 32)) & M32
    v = {"a%d" % i: (aval >> (32 * i)) & M32 for i in range(8)}

    def val(s):
        if isinstance(s, int):
            return s
        if isinstance(s, tuple):
            return inv if s[0] == "inv" else (p >> (32 * s[1])) & M32
        return v[s]

    for b in blocks:
        cf = None   # undefined at the start of a statement
        for t in b.ins:
            op, dst = t[0], t[1]
            s = [val(z) for z in t[2:]]
            parts = op.split(".")
            base = parts[0]
            cc_out = parts[-1] == "cc"
            if base == "mov":
                r = s[0]
                out_c = None
            elif base == "shr31":
                r = s[0] >> 31
                out_c = None
            elif base == "shl1":
                r = (s[0] << 1) & M32
                out_c = None
            elif base == "addw":
                r = (s[0] + s[1]) & M32
                out_c = None
            elif base == "shf":
                r = ((s[1] << 1) | (s[0] >> 31)) & M32
                out_c = None
            elif base in ("mul",):
                pr = s[0] * s[1]
                r = (pr & M32) if parts[1] == "lo" else (pr >> 32)
                out_c = None
            elif base in ("mad", "madc"):
                pr = s[0] * s[1]
                r = ((pr & M32) if parts[1] == "lo" else (pr >> 32)) + s[2]
                if base == "madc":
                    assert cf is not None, ("carry read but not set", t)
                    r += cf
                out_c = r >> 32
                r &= M32
            elif base in ("add", "addc"):
                r = s[0] + s[1]
                if base == "addc":
                    assert cf is not None, ("carry read but not set", t)
                    r += cf
                out_c = r >> 32
                r &= M32
            else:
                raise ValueError(op)
            if cc_out:
                cf = out_c
            elif out_c:
                raise AssertionError(("carry lost", t))   # an overflow nobody consumes = wrong result
            if not cc_out and base in ("madc", "addc"):
                cf = None   # consumed; the next reader must set its own
            v[dst] = r
    t = sum(v["t%d" % j] << (32 * j) for j in range(8))
    return t


def check(n_random=20000, n_extreme=2000):
    blocks = build()
    n_wide = 0
    n_other = 0
    for b in blocks:
        for t in b.ins:
            if t[0].startswith(("mad", "mul")) and not (isinstance(t[3], tuple) and t[3][0] == "inv"):
                n_wide += 1
            else:
                n_other += 1
    print("instructions: %d half-MACs (= %d wide MACs), %d others" % (n_wide, n_wide // 2, n_other))
    rng = random.Random(7)
    for p in (Q, R_):
        rinv = pow(1 << 256, -1, p)
        cases = [0, 1, 2, p - 1, p - 2, (1 << 253) - 1, (1 << 224) - 1, M32, p >> 1]
        cases += [sum(M32 << (32 * i) for i in range(8)) % p]
        cases += [rng.randrange(p) for _ in range(n_random)]
        # operands with extreme limbs: all-ones limbs force every carry
        for _ in range(n_extreme):
            a = 0
            for i in range(8):
                a |= rng.choice((0, M32, 1, 0x80000000, rng.randrange(1 << 32))) << (32 * i)
            cases.append(a % p)
        for a in cases:
            t = emulate(blocks, a, p)
            want = a * a * rinv % p
            assert t < 2 * p and t % p == want, (hex(a), hex(t), hex(want))
    print("fp_sqr instruction list: %d operands per field ok (result < 2p, congruent to a^2 / R)" % len(cases))


# ------------------------------------------------------------------------------------------------------------------
def emit(blocks=None, inputs=("a",)):
    if blocks is None:
        blocks = build()
    out = []
    declared = set("%s%d" % (nm, i) for nm in inputs for i in range(8))
    decl = []
    for b in blocks:
        for t in b.ins:
            if t[1] not in declared:
                declared.add(t[1])
                decl.append(t[1])
    for nm in inputs:
        out.append("    const uint32_t " + ", ".join("%s%d = %s.l[%d]" % (nm, i, nm, i) for i in range(8)) + ";")
    line = "    uint32_t "
    cur = line
    for i, d in enumerate(decl):
        piece = d + (", " if i + 1 < len(decl) else ";")
        if len(cur) + len(piece) > 118:
            out.append(cur.rstrip())
            cur = "        "
        cur += piece
    out.append(cur)
    for b in blocks:
        ops = []      # (constraint, c-expression)
        index = {}

        def ref(s, write=False):
            if isinstance(s, int):
                return str(s)
            key = s
            if key not in index:
                index[key] = len(ops)
                if isinstance(s, tuple):
                    ops.append(["r", "P::INV" if s[0] == "inv" else "P::mod(%d)" % s[1], False, False])
                else:
                    ops.append(["r", s, False, False])
            return "%%%d" % index[key]

        # first pass: which variables are written, and whether they are read before their first write
        wr, rd_first = [], set()
        seen_w = set()
        for t in b.ins:
            for s in t[2:]:
                if isinstance(s, str) and s not in seen_w:
                    rd_first.add(s)
            if t[1] not in seen_w:
                seen_w.add(t[1])
                wr.append(t[1])
        # outputs first, then inputs (PTX operand numbering)
        for wv in wr:
            index[wv] = len(ops)
            ops.append(["+r" if wv in rd_first else "=r", wv, True, wv in rd_first])
        lines = []
        for t in b.ins:
            op = t[0]
            name = {"shf.l": "shf.l.wrap.b32", "mov": "mov.b32"}.get(op, op + ".u32")
            if op == "addw":
                name = "add.u32"
            if op == "and":
                name = "and.b32"
            if op == "neg":
                name = "neg.s32"
            if op == "shr31":
                name, args = "shr.u32", [ref(t[1]), ref(t[2]), "31"]
            elif op == "shl1":
                name, args = "shl.b32", [ref(t[1]), ref(t[2]), "1"]
            elif op == "shf.l":
                args = [ref(t[1]), ref(t[2]), ref(t[3]), "1"]
            else:
                args = [ref(t[1])] + [ref(s) for s in t[2:]]
            lines.append("%s %s;" % (name, ", ".join(args)))
        outs = [o for o in ops if o[2]]
        ins = [o for o in ops if not o[2]]
        text = '    asm("' + '\\n\\t"\n        "'.join(lines) + '"\n'
        text += "        : " + ", ".join('"%s"(%s)' % (o[0], o[1]) for o in outs) + "\n"
        text += "        : " + ", ".join('"r"(%s)' % o[1] for o in ins) + ");"
        out.append(text)
    out.append("    uint32_t t[8] = {t0, t1, t2, t3, t4, t5, t6, t7};")
    out.append("    return fp_final_sub<P>(t);")
    return "\n".join(out)


BEGIN = "// ---- GENERATED by tools/gen_fp_sqr.py (do not edit by hand) ----"
END = "// ---- end of generated code ----"


BEGIN_K = "// ---- GENERATED by tools/gen_fp_sqr.py: Karatsuba product (do not edit by hand) ----"
END_K = "// ---- end of the generated Karatsuba product ----"


def update():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kzg_grandsums_study_b200", "csrc", "field.cuh")
    s = open(path).read()
    i, j = s.index(BEGIN), s.index(END)
    s = s[:i] + BEGIN + "\n" + emit() + "\n    " + s[j:]
    if BEGIN_K in s:
        i, j = s.index(BEGIN_K), s.index(END_K)
        s = s[:i] + BEGIN_K + "\n" + emit(build_mulk(), inputs=("a", "b")) + "\n    " + s[j:]
    open(path, "w").write(s)
    print("updated", path)


# ------------------------------------------------------------------------------------------------------------------
# fp_mul2(a, b, c, d) = (a b + c d) / R mod p with ONE interleaved reduction (192 wide MACs instead of 256): the rows of
# fp_mul with a second product row added before the reduction row.  The C++ is hand-written from the helpers of fp_mul
# (cmad4 / cmad4_top / shift_mad4); this emulation of the same helper sequence checks the carry bounds (a "carry lost"
# assertion = an overflow the code would drop) and that the result is < 3p (two conditional subtractions).
def build_mul2():
    blocks = []
    a = ["a%d" % i for i in range(8)]
    bb = ["b%d" % i for i in range(8)]
    c = ["c%d" % i for i in range(8)]
    d = ["d%d" % i for i in range(8)]
    x = ["x%d" % k for k in range(8)]
    y = ["y%d" % k for k in range(8)]
    mod = [("mod", k) for k in range(8)]

    def cmad4(acc, xs, yv):          # acc[0..7] += {xs} * y, carry out dropped
        b = Block()
        for n in range(4):
            b.add("mad.lo.cc" if n == 0 else "madc.lo.cc", acc[2 * n], xs[n], yv, acc[2 * n])
            b.add("madc.hi" + (".cc" if n < 3 else ""), acc[2 * n + 1], xs[n], yv, acc[2 * n + 1])
        blocks.append(b)

    def cmad4_top(acc, xs, yv, top):
        b = Block()
        for n in range(4):
            b.add("mad.lo.cc" if n == 0 else "madc.lo.cc", acc[2 * n], xs[n], yv, acc[2 * n])
            b.add("madc.hi.cc", acc[2 * n + 1], xs[n], yv, acc[2 * n + 1])
        b.add("addc", top, top, 0)
        blocks.append(b)

    def shift_mad4(lo0, cw, hi, xs, yv):
        b = Block()
        b.add("add.cc", lo0, lo0, cw)
        src = [hi[2], hi[3], hi[4], hi[5], hi[6], hi[7], 0, 0]
        for n in range(4):
            b.add("madc.lo.cc", hi[2 * n], xs[n], yv, src[2 * n])
            b.add("madc.hi" + (".cc" if n < 3 else ""), hi[2 * n + 1], xs[n], yv, src[2 * n + 1])
        blocks.append(b)

    ev = lambda v: [v[0], v[2], v[4], v[6]]
    od = lambda v: [v[1], v[3], v[5], v[7]]
    b = Block()
    for n in range(4):
        b.add("mul.lo", x[2 * n], ev(a)[n], bb[0])
        b.add("mul.hi", x[2 * n + 1], ev(a)[n], bb[0])
    for n in range(4):
        b.add("mul.lo", y[2 * n], od(a)[n], bb[0])
        b.add("mul.hi", y[2 * n + 1], od(a)[n], bb[0])
    blocks.append(b)
    cmad4(y, od(c), d[0])
    cmad4_top(x, ev(c), d[0], y[7])
    b = Block()
    b.add("mul.lo", "m", x[0], ("inv",))
    blocks.append(b)
    cmad4(y, od(mod), "m")
    cmad4_top(x, ev(mod), "m", y[7])
    for i in range(1, 8):
        lo, hi = (y, x) if i & 1 else (x, y)
        shift_mad4(lo[0], hi[1], hi, od(a), bb[i])
        cmad4_top(lo, ev(a), bb[i], hi[7])
        cmad4(hi, od(c), d[i])
        cmad4_top(lo, ev(c), d[i], hi[7])
        b = Block()
        b.add("mul.lo", "m", lo[0], ("inv",))
        blocks.append(b)
        cmad4(hi, od(mod), "m")
        cmad4_top(lo, ev(mod), "m", hi[7])
    lo, hi = y, x
    b = Block()
    for j in range(8):
        op = "add.cc" if j == 0 else ("addc.cc" if j < 7 else "addc")
        b.add(op, "t%d" % j, hi[j], lo[j + 1] if j < 7 else 0)
    blocks.append(b)
    return blocks


def emulate_named(blocks, vals, p):
    """like emulate(), inputs given as {prefix: integer} for 8-limb operands"""
    inv = (-pow(p, -1, 1 << 32)) & M32
    v = {}
    for k, val_ in vals.items():
        for i in range(8):
            v["%s%d" % (k, i)] = (val_ >> (32 * i)) & M32
    fake = [Block()]
    # reuse emulate()'s interpreter by pre-seeding: emulate() seeds a0..a7 only, so run the loop here
    def val(s):
        if isinstance(s, int):
            return s
        if isinstance(s, tuple):
            return inv if s[0] == "inv" else (p >> (32 * s[1])) & M32
        return v[s]
    for b in blocks:
        cf = None
        for t in b.ins:
            op, dst = t[0], t[1]
            s = [val(z) for z in t[2:]]
            parts = op.split(".")
            base = parts[0]
            cc_out = parts[-1] == "cc"
            if base == "mul":
                pr = s[0] * s[1]
                r = (pr & M32) if parts[1] == "lo" else (pr >> 32)
                out_c = None
            elif base in ("mad", "madc"):
                pr = s[0] * s[1]
                r = ((pr & M32) if parts[1] == "lo" else (pr >> 32)) + s[2]
                if base == "madc":
                    assert cf is not None
                    r += cf
                out_c = r >> 32
                r &= M32
            elif base in ("add", "addc"):
                r = s[0] + s[1]
                if base == "addc":
                    assert cf is not None
                    r += cf
                out_c = r >> 32
                r &= M32
            elif base in ("sub", "subc"):
                r = s[0] - s[1]
                if base == "subc":
                    assert cf is not None
                    r -= cf
                out_c = 1 if r < 0 else 0   # the borrow
                r &= M32
            elif base == "and":
                r = s[0] & s[1]
                out_c = None
            elif base == "neg":
                r = (-s[0]) & M32
                out_c = None
            elif base == "mov":
                r = s[0]
                out_c = None
            elif base == "addw":
                r = (s[0] + s[1]) & M32
                out_c = None
            else:
                raise ValueError(op)
            if cc_out:
                cf = out_c
            elif out_c:
                raise AssertionError(("carry lost", t))
            if not cc_out and base in ("madc", "addc", "subc"):
                cf = None
            v[dst] = r
    return sum(v["t%d" % j] << (32 * j) for j in range(8))


def check_mul2(n_random=20000, n_extreme=4000):
    blocks = build_mul2()
    rng = random.Random(11)
    worst = 0
    for p in (Q, R_):
        rinv = pow(1 << 256, -1, p)
        edge = [0, 1, p - 1, p - 2, (1 << 253) - 1, p >> 1, M32, sum(M32 << (32 * i) for i in range(8)) % p]
        cases = [(a_, b_, c_, d_) for a_ in edge for b_ in edge for c_ in edge for d_ in edge]
        cases += [tuple(rng.randrange(p) for _ in range(4)) for _ in range(n_random)]
        for _ in range(n_extreme):
            ops = []
            for _k in range(4):
                a_ = 0
                for i in range(8):
                    a_ |= rng.choice((0, M32, 1, 0x80000000, rng.randrange(1 << 32))) << (32 * i)
                ops.append(a_ % p)
            cases.append(tuple(ops))
        for (a_, b_, c_, d_) in cases:
            t = emulate_named(blocks, {"a": a_, "b": b_, "c": c_, "d": d_}, p)
            want = (a_ * b_ + c_ * d_) * rinv % p
            assert t % p == want and t < 3 * p, (hex(a_), hex(b_), hex(c_), hex(d_), hex(t))
            worst = max(worst, t * 1000 // p)
    print("fp_mul2 helper sequence: %d operand tuples per field ok; largest result %.3f p (< 3 p)" % (len(cases), worst / 1000))



# ------------------------------------------------------------------------------------------------------------------
# fp_mulk(a, b): Karatsuba 8 x 8 limbs (three 4 x 4 products = 48 wide MACs) + the separated reduction above (64):
# 112 wide MACs instead of 128, paid for with ~70 more additions on the (mostly idle) ALU pipe.
#   a b = z0 + ((aL + aH)(bL + bH) - z0 - z2) 2^128 + z2 2^256
# Every 4 x 4 product uses the even / odd accumulator pair (E: even limb positions, O: odd) so that a row is two carry
# chains of two word pairs; a chain's carry lands in a word that holds nothing but carries so far.
def mac_chain(blocks, written, acc, base, prods, carry_to):
    """acc[base + 2 n], acc[base + 2 n + 1] += x_n * y_n for (x_n, y_n) in prods, one carry chain; carry -> acc[carry_to]"""
    b = Block()
    nocarry = True
    for n, (xv, yv) in enumerate(prods):
        for k, part in ((0, "lo"), (1, "hi")):
            word = acc[base + 2 * n + k]
            addend = word if word in written else 0
            last = n == len(prods) - 1 and part == "hi"
            cc_out = not (last and carry_to is None)
            if nocarry:
                if addend == 0:
                    b.add("mul." + part, word, xv, yv)
                else:
                    b.add("mad.%s%s" % (part, ".cc" if cc_out else ""), word, xv, yv, addend)
                    nocarry = False
            else:
                b.add("madc.%s%s" % (part, ".cc" if cc_out else ""), word, xv, yv, addend)
            written.add(word)
    if carry_to is not None:
        w = acc[carry_to]
        assert w not in written and not nocarry
        b.add("addc", w, 0, 0)
        written.add(w)
    blocks.append(b)


def subprod(blocks, x, y, name):
    """-> names of the 8 words of x[0..3] * y[0..3]"""
    E = ["%se%d" % (name, k) for k in range(8)]
    O = ["%so%d" % (name, k) for k in range(7)]     # O[k] = word position k + 1
    w = set()
    mac_chain(blocks, w, E, 0, [(x[0], y[0]), (x[2], y[0])], None)
    mac_chain(blocks, w, O, 0, [(x[1], y[0]), (x[3], y[0])], None)
    mac_chain(blocks, w, O, 0, [(x[0], y[1]), (x[2], y[1])], 4)
    mac_chain(blocks, w, E, 2, [(x[1], y[1]), (x[3], y[1])], None)
    mac_chain(blocks, w, E, 2, [(x[0], y[2]), (x[2], y[2])], 6)
    mac_chain(blocks, w, O, 2, [(x[1], y[2]), (x[3], y[2])], None)
    mac_chain(blocks, w, O, 2, [(x[0], y[3]), (x[2], y[3])], 6)
    mac_chain(blocks, w, E, 4, [(x[1], y[3]), (x[3], y[3])], None)
    assert w == set(E) | set(O)
    b = Block()
    for k in range(1, 8):
        b.add("add.cc" if k == 1 else ("addc.cc" if k < 7 else "addc"), E[k], E[k], O[k - 1])
    blocks.append(b)
    return E


def build_mulk():
    blocks = []
    a = ["a%d" % i for i in range(8)]
    bb = ["b%d" % i for i in range(8)]
    z0 = subprod(blocks, a[:4], bb[:4], "p")
    z2 = subprod(blocks, a[4:], bb[4:], "q")
    sa = ["sa%d" % i for i in range(4)]
    sb = ["sb%d" % i for i in range(4)]
    for sv, v, cv in ((sa, a, "ca"), (sb, bb, "cb")):
        b = Block()
        for i in range(4):
            b.add("add.cc" if i == 0 else "addc.cc", sv[i], v[i], v[4 + i])
        b.add("addc", cv, 0, 0)
        blocks.append(b)
    zm = subprod(blocks, sa, sb, "m") + ["zm8"]
    b = Block()
    b.add("and", "zm8", "ca", "cb")
    b.add("neg", "na", "ca")
    b.add("neg", "nb", "cb")
    for i in range(4):
        b.add("and", "ta%d" % i, sb[i], "na")
        b.add("and", "tb%d" % i, sa[i], "nb")
    blocks.append(b)
    for tv in ("ta", "tb"):     # zm += (ca sb + cb sa) 2^128
        b = Block()
        for i in range(4):
            b.add("add.cc" if i == 0 else "addc.cc", zm[4 + i], zm[4 + i], "%s%d" % (tv, i))
        b.add("addc", "zm8", "zm8", 0)
        blocks.append(b)
    for z in (z0, z2):          # zm -= z0, zm -= z2 (never negative)
        b = Block()
        for i in range(8):
            b.add("sub.cc" if i == 0 else "subc.cc", zm[i], zm[i], z[i])
        b.add("subc", "zm8", "zm8", 0)
        blocks.append(b)
    T = z0 + z2
    b = Block()
    for i in range(9):
        b.add("add.cc" if i == 0 else "addc.cc", T[4 + i], T[4 + i], zm[i])
    b.add("addc.cc", T[13], T[13], 0)
    b.add("addc.cc", T[14], T[14], 0)
    b.add("addc", T[15], T[15], 0)
    blocks.append(b)
    append_redc(blocks, T)
    return blocks


def check_mulk(n_random=20000, n_extreme=6000):
    blocks = build_mulk()
    n_wide = sum(1 for b in blocks for t in b.ins
                 if t[0].startswith(("mad", "mul")) and not (isinstance(t[3], tuple) and t[3][0] == "inv"))
    n_all = sum(len(b.ins) for b in blocks)
    n_mov = sum(1 for b in blocks for t in b.ins if t[0] == "mov")
    print("fp_mulk: %d wide MACs, %d other instructions (+ %d register moves)" % (n_wide // 2, n_all - n_wide - n_mov, n_mov))
    rng = random.Random(13)
    for p in (Q, R_):
        rinv = pow(1 << 256, -1, p)
        edge = [0, 1, p - 1, p - 2, (1 << 253) - 1, p >> 1, M32, (1 << 128) - 1, ((1 << 128) - 1) << 128 & ((1 << 253) - 1),
                sum(M32 << (32 * i) for i in range(8)) % p, (1 << 128), (1 << 127) | (1 << 253) % p]
        cases = [(x_, y_) for x_ in edge for y_ in edge]
        cases += [(rng.randrange(p), rng.randrange(p)) for _ in range(n_random)]
        for _ in range(n_extreme):
            ops = []
            for _k in range(2):
                a_ = 0
                for i in range(8):
                    a_ |= rng.choice((0, M32, 1, 0x80000000, 0xFFFFFFFE, rng.randrange(1 << 32))) << (32 * i)
                ops.append(a_ % p)
            cases.append(tuple(ops))
        for (a_, b_) in cases:
            t = emulate_named(blocks, {"a": a_, "b": b_}, p)
            assert t % p == a_ * b_ * rinv % p and t < 2 * p, (hex(a_), hex(b_), hex(t))
    print("fp_mulk instruction list: %d operand pairs per field ok" % len(cases))


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else "check"
    if cmd == "check":
        check()
        check_mul2()
        check_mulk()
    elif cmd == "check2":
        check_mul2()
    elif cmd == "checkk":
        check_mulk()
    elif cmd == "emit":
        print(emit())
    elif cmd == "emitk":
        print(emit(build_mulk(), inputs=("a", "b")))
    elif cmd == "update":
        check()
        check_mulk()
        update()
