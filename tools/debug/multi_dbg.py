import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.py import bn254 as bn, inputs
from kzg_grandsums_study_b200.curve import getCurveFromName
from kzg_grandsums_study_b200._lib import as_ptr
R = bn.R
curve = getCurveFromName("bn128")
tau = inputs.tau_from_seed(1001)
n = 3000
def closed(coeffs):
    return bn.g1_to_bytes(bn.g1_mul_gen(sum(c * pow(tau, i, R) for i, c in enumerate(coeffs)) % R))
polys = [inputs.random_column(31, 1500), inputs.random_column(32, 3000), [0] * 700, [7] * 2048, [], [R - 1] * 33, inputs.random_column(33, 1)]
want = [closed(p) for p in polys]
for table_c in (5, 11):
    srs = C.c_void_p()
    curve.check(curve.lib.kzg_srs_generate(curve.ctx, as_ptr(tau.to_bytes(32, "little")), n, C.byref(srs)))
    curve.check(curve.lib.kzg_srs_precompute(curve.ctx, srs, table_c))
    bufs = [curve.to_device(bn.fr_vec_to_mont_bytes(p)) for p in polys]
    for rounds in (0, 1, 2, 5):
        curve.set_option("aff_rounds", rounds)
        for merge in (1, 0):
            curve.set_option("msm_merge", merge)
            for rep in range(3):
                k = len(bufs)
                handles = (C.c_void_p * k)(*[b.handle for b in bufs])
                out = bytearray(64 * k)
                curve.check(curve.lib.kzg_commit_many(curve.ctx, srs, handles, k, as_ptr(out)))
                got = [bytes(out[64 * i:64 * i + 64]) for i in range(k)]
                bad = [i for i in range(k) if got[i] != want[i]]
                print("c=%d rounds=%d merge=%d rep=%d bad=%s" % (table_c, rounds, merge, rep, bad), flush=True)
        # singles
        for i, b in enumerate(bufs):
            o = bytearray(64)
            curve.check(curve.lib.kzg_commit(curve.ctx, srs, b.handle, as_ptr(o)))
            if bytes(o) != want[i]:
                print("  single commit c=%d rounds=%d job %d BAD" % (table_c, rounds, i), flush=True)
    curve.set_option("aff_rounds", -1)
    curve.lib.kzg_srs_free(curve.ctx, srs)
