"""A few NTTs of one size (for ncu captures):  python tools/ntt_once.py LOG_N [REPS]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kzg_grandsums_study_b200 import synthetic  # noqa: E402
from kzg_grandsums_study_b200.curve import Curve  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
_stream = torch.cuda.Stream()
torch.cuda.set_stream(_stream)
curve = Curve(0, _stream.cuda_stream)
n = 1 << log_n
x = curve.to_device(synthetic.random_fr_std(1, n).tobytes())
y = curve.alloc(n)
for i in range(reps):
    curve.check(curve.lib.kzg_fr_ntt(curve.ctx, x.handle, y.handle, 0))
    curve.check(curve.lib.kzg_fr_ntt(curve.ctx, y.handle, y.handle, 1))
torch.cuda.synchronize()
print("ntt 2^%d x %d fwd+inv ok: %s" % (log_n, reps, y.slice(0, 32) == x.slice(0, 32)))
