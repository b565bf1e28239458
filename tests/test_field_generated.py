"""The generated field routines of csrc/field.cuh (dedicated Montgomery squaring: 100 wide MACs instead of 128) and the
helper sequence of fp_mul2 (two products under one reduction): tools/gen_fp_sqr.py holds their instruction lists, emulates
them word by word with an explicit carry flag -- every carry a statement drops is asserted zero -- and prints the CUDA text.
Here, on the CPU: the emulation against a^2 / R and (a b + c d) / R mod p for both fields (random, edge and extreme-limb
operands), and the block between the GENERATED markers of field.cuh is exactly what the generator prints.  On the device
kzg_selftest() compares the compiled code with fp_mul_portable (tests/test_gpu_primitives.py, smoke()).
These replace ffjavascript's Fr / F1 square and mul inside every bulk call of the path (reference src/polynomial/polynomial.js:1106-1115)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gen():
    spec = importlib.util.spec_from_file_location("gen_fp_sqr", os.path.join(ROOT, "tools", "gen_fp_sqr.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_squaring_instruction_list_emulates_to_the_montgomery_square():
    g = _gen()
    g.check(n_random=1500, n_extreme=500)
    blocks = g.build()
    wide = sum(1 for b in blocks for t in b.ins if t[0].startswith(("mad", "mul")) and not (isinstance(t[3], tuple) and t[3][0] == "inv"))
    assert wide == 200            # 100 wide MACs (lo / hi halves)


def test_two_products_under_one_reduction_stay_below_2p():
    g = _gen()
    g.check_mul2(n_random=1000, n_extreme=500)


def test_field_cuh_holds_exactly_the_generated_squaring():
    g = _gen()
    text = open(os.path.join(ROOT, "kzg_grandsums_study_b200", "csrc", "field.cuh")).read()
    i, j = text.index(g.BEGIN), text.index(g.END)
    body = text[i + len(g.BEGIN):j].strip("\n").rstrip()
    assert body == g.emit().rstrip(), "field.cuh: the generated block differs from tools/gen_fp_sqr.py emit (run `update`)"
