"""Device timing of the bulk Fr primitives (CUDA events on the context's stream):
   python tools/ops_bench.py [LOG_N ...]      -> NTT / iNTT / batch inverse / grand-sum build / evaluate / to_mont"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kzg_grandsums_study_b200 import synthetic  # noqa: E402
from kzg_grandsums_study_b200._lib import as_ptr  # noqa: E402
from kzg_grandsums_study_b200.curve import Curve  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [16, 20, 22, 24]
_stream = torch.cuda.Stream()
torch.cuda.set_stream(_stream)
curve = Curve(0, _stream.cuda_stream)
lib, ctx = curve.lib, curve.ctx
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=5):
    fn()
    best = 1e9
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


for lg in sizes:
    n = 1 << lg
    x = curve.to_device(synthetic.random_fr_std(lg, n).tobytes())
    y = curve.alloc(n)
    gamma = (12345).to_bytes(32, "little")
    res = {}
    res["ntt"] = timed(lambda: curve.check(lib.kzg_fr_ntt(ctx, x.handle, y.handle, 0)))
    res["intt"] = timed(lambda: curve.check(lib.kzg_fr_ntt(ctx, x.handle, y.handle, 1)))
    res["batch_inverse"] = timed(lambda: curve.check(lib.kzg_fr_batch_inverse(ctx, x.handle, y.handle)))
    res["to_mont"] = timed(lambda: curve.check(lib.kzg_fr_to_mont(ctx, x.handle, y.handle)))
    out = bytearray(32)
    res["evaluate"] = timed(lambda: curve.check(lib.kzg_poly_evaluate(ctx, x.handle, as_ptr(gamma), as_ptr(out))))

    def build():
        h = C.c_void_p()
        rc = lib.kzg_grandproduct_build(ctx, x.handle, x.handle, None, None, as_ptr(gamma), C.byref(h))
        curve.check(rc)
        lib.kzg_buf_free(ctx, h)
    res["grandproduct_build"] = timed(build)
    gb = 64.0 * n / 1e9
    print("n=2^%d: " % lg + ", ".join("%s %.3f ms (%.0f GB/s alg.)" % (k, v, gb / (v * 1e-3)) for k, v in res.items()))
