"""Per-phase device times of the SRS (window-table) MSM, CUDA events on the launching stream:
python tools/msm_phases.py [LOG_N ...]   ->  sort (digits/scan/scatter) | accumulate | reduce | finish | whole call
ROUNDS="0,1,2" CHUNKS="1,2" sweep the number of batched-affine rounds / chunks per round for every size (-1 = library default)"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kzg_grandsums_study_b200 import synthetic  # noqa: E402
from kzg_grandsums_study_b200._lib import as_ptr  # noqa: E402
from kzg_grandsums_study_b200.curve import Curve  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [16, 18, 20, 21, 22, 24]
_stream = torch.cuda.Stream()
torch.cuda.set_stream(_stream)
curve = Curve(0, _stream.cuda_stream)
lib, ctx = curve.lib, curve.ctx
tau = synthetic.tau_from_seed(1001)
TAGS = (("sort", 2), ("affine", 5), ("accumulate", 0), ("reduce", 3), ("finish", 4))
for log_n in sizes:
    n = 1 << log_n
    srs = C.c_void_p()
    curve.check(lib.kzg_srs_generate(ctx, as_ptr(tau.to_bytes(32, "little")), n, C.byref(srs)))
    curve.check(lib.kzg_srs_precompute(ctx, srs, int(os.environ.get("TABLE_C", "0"))))
    scal = curve.to_device(synthetic.random_fr_std(6, n).tobytes())
    out = bytearray(64)
    wb, nw = C.c_uint32(), C.c_uint32()
    curve.check(lib.kzg_msm_geometry(ctx, srs, n, 0, C.byref(wb), C.byref(nw)))
    for rounds in [int(x) for x in os.environ.get("ROUNDS", "-1").split(",")]:
        for chunks in [int(x) for x in os.environ.get("CHUNKS", "-1").split(",")]:
            curve.set_option("aff_rounds", rounds)
            curve.set_option("aff_chunks", chunks)
            for _ in range(3):
                curve.check(lib.kzg_srs_msm(ctx, srs, 0, scal.handle, n, as_ptr(out)))
            reps = 5
            curve.check(lib.kzg_ctx_kernel_time(ctx, 0, 1, None, None))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                curve.check(lib.kzg_srs_msm(ctx, srs, 0, scal.handle, n, as_ptr(out)))
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3 / reps
            parts = []
            for name, tag in TAGS:
                ms, cnt = C.c_double(), C.c_uint64()
                curve.check(lib.kzg_ctx_kernel_time(ctx, tag, 0, C.byref(ms), C.byref(cnt)))
                parts.append("%s %.3f" % (name, ms.value / reps))
            curve.check(lib.kzg_ctx_kernel_time(ctx, 0, -1, None, None))
            print("msm 2^%d c=%d windows=%d rounds=%d chunks=%d: %s | call %.3f ms" % (
                log_n, wb.value, nw.value, rounds, chunks, " | ".join(parts), wall), flush=True)
    del scal
    lib.kzg_srs_free(ctx, srs)
