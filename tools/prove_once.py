"""One grand-sum (or grand-product) proof at n = 2^LOG_N with per-round host timing:
   python tools/prove_once.py LOG_N [gs|gp] [REPS]"""
import ctypes as C
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kzg_grandsums_study_b200 import _lib, synthetic, curve as curve_mod  # noqa: E402
from kzg_grandsums_study_b200._lib import as_ptr  # noqa: E402
from kzg_grandsums_study_b200.grandsum import mset_eq_kzg_grandsum_prover  # noqa: E402
from kzg_grandsums_study_b200.grandproduct import mset_eq_kzg_grandproduct_prover  # noqa: E402
from kzg_grandsums_study_b200.polynomial import Evaluations  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
kind = sys.argv[2] if len(sys.argv) > 2 else "gs"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n = 1 << log_n
curve = curve_mod.getCurveFromName("bn128")
lib, ctx = curve.lib, curve.ctx
tau = synthetic.tau_from_seed(1001)
f = synthetic.random_fr_std(4, n)
t = f[synthetic.permutation(4, n)]
import numpy as np  # noqa: E402
fb = torch.from_numpy(f.view(np.uint8).reshape(-1).copy()).pin_memory()
tb = torch.from_numpy(np.ascontiguousarray(t).view(np.uint8).reshape(-1).copy()).pin_memory()
prover = mset_eq_kzg_grandsum_prover if kind == "gs" else mset_eq_kzg_grandproduct_prover

# per-call timing of the C ABI rounds
timings = {}
for name in ("kzg_prover_create", "kzg_prover_round1", "kzg_prover_round2", "kzg_prover_round3", "kzg_prover_round4",
             "kzg_prover_round5", "kzg_prover_destroy", "kzg_srs_load_ptau"):
    orig = getattr(lib, name)

    def wrap(*a, _orig=orig, _name=name):
        t0 = time.perf_counter()
        r = _orig(*a)
        timings.setdefault(_name, []).append((time.perf_counter() - t0) * 1e3)
        return r
    setattr(lib, name, wrap)

with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "p.ptau")
    srs = C.c_void_p()
    curve.check(lib.kzg_srs_generate(ctx, as_ptr(tau.to_bytes(32, "little")), 2 * n, C.byref(srs)))
    z = bytes(128)
    curve.check(lib.kzg_srs_write_ptau(ctx, srs, log_n, as_ptr(z), as_ptr(z), path.encode()))
    lib.kzg_srs_free(ctx, srs)
    for i in range(reps):
        timings.clear()
        torch.cuda.synchronize()
        l0 = curve.launch_count()
        t0 = time.perf_counter()
        proof = prover(path, Evaluations(fb, curve), Evaluations(tb, curve))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        print("%s prove n=2^%d rep %d: %.2f ms, %d launches; %s" % (
            kind, log_n, i, dt, curve.launch_count() - l0,
            ", ".join("%s %.2f" % (k.replace("kzg_prover_", "").replace("kzg_", ""), sum(v)) for k, v in timings.items())))
