"""The C-level multi-GPU MSM (kzg_mgpu_*, one process, one context and one host thread per device) at 2^LOG_N points:
   python tools/mgpu_bench.py [LOG_N] [DEVICES e.g. 0,1,2,3]      -> resident and host-scalar (e2e) ms, Mpts/s"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from kzg_grandsums_study_b200 import _lib, synthetic  # noqa: E402
from kzg_grandsums_study_b200._lib import as_ptr  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
devices = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else list(range(torch.cuda.device_count()))
lib = _lib.load()
h = C.c_void_p()
assert lib.kzg_mgpu_create((C.c_int * len(devices))(*devices), len(devices), C.byref(h)) == 0
n = 1 << log_n
tau = synthetic.tau_from_seed(1001)
t0 = time.perf_counter()
assert lib.kzg_mgpu_srs_generate(h, as_ptr(tau.to_bytes(32, "little")), n) == 0, lib.kzg_mgpu_last_error(h)
print("SRS shards + window tables on %d device(s): %.1f ms" % (len(devices), (time.perf_counter() - t0) * 1e3))
scal = torch.from_numpy(synthetic.random_fr_std(6, n).view(np.int64).copy()).pin_memory()
out = bytearray(64)
assert lib.kzg_mgpu_scalars_upload(h, as_ptr(scal), n) == 0


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t) * 1e3)
    return min(ts), sorted(ts)[len(ts) // 2]


best, med = timed(lambda: lib.kzg_mgpu_srs_msm(h, as_ptr(out)))
print("mgpu msm 2^%d on devices %s: resident best %.3f ms (%.0f Mpts/s), median %.3f ms" % (log_n, devices, best, n / best / 1e3, med))
res = bytes(out)
best, med = timed(lambda: lib.kzg_mgpu_srs_msm_host(h, as_ptr(scal), n, as_ptr(out)))
print("mgpu msm 2^%d on devices %s: e2e (pinned host scalars) best %.3f ms (%.0f Mpts/s), median %.3f ms; same point: %s" % (
    log_n, devices, best, n / best / 1e3, med, bytes(out) == res))
lib.kzg_mgpu_destroy(h)
