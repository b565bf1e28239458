"""Run a few MSMs of one size (for ncu captures and quick timing):  python tools/msm_once.py LOG_N [WINDOW] [REPS]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kzg_grandsums_study_b200 import synthetic  # noqa: E402
from kzg_grandsums_study_b200._lib import as_ptr  # noqa: E402
from kzg_grandsums_study_b200.curve import Curve  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
window = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n = 1 << log_n
_stream = torch.cuda.Stream()
torch.cuda.set_stream(_stream)
curve = Curve(0, _stream.cuda_stream)
lib, ctx = curve.lib, curve.ctx
tau = synthetic.tau_from_seed(1001)
srs = C.c_void_p()
curve.check(lib.kzg_srs_generate(ctx, as_ptr(tau.to_bytes(32, "little")), n, C.byref(srs)))
scal = curve.to_device(synthetic.random_fr_std(6, n).tobytes())
if window > 0:
    curve.check(lib.kzg_msm_set_window(ctx, window))       # per-window (table-less) path
else:
    curve.check(lib.kzg_srs_precompute(ctx, srs, -window))  # window table; 0 = auto, -c = table with c bits
out = bytearray(64)
for i in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    curve.check(lib.kzg_srs_msm(ctx, srs, 0, scal.handle, n, as_ptr(out)))
    torch.cuda.synchronize()
    print("msm 2^%d window %d rep %d: %.3f ms  -> %s" % (log_n, window, i, (time.perf_counter() - t0) * 1e3, bytes(out[:8]).hex()))
