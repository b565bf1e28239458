#!/bin/bash
# session 2, call 6: Karatsuba product (fp_mulk, 112 wide MACs) against the interleaved CIOS product (128): A/B of two builds
mkdir -p gpurun_out
for lib in "" kzg_grandsums_study_b200/variants/libkzgb200_k.so; do
echo "=== lib: ${lib:-default}"
KZGB200_LIB=$lib timeout 120 python - <<'PY'
import ctypes as C
from kzg_grandsums_study_b200.curve import getCurveFromName
c = getCurveFromName("bn128")
c.check(c.lib.kzg_selftest(c.ctx, 1 << 16))
print("selftest ok")
for name in ("kzg_bench_imad_peak", "kzg_bench_modmul_peak"):
    v = C.c_double()
    c.check(getattr(c.lib, name)(c.ctx, 200, C.byref(v)))
    print(name, "%.3f T MAC/s" % (v.value / 1e12), "= %.1f G products/s" % (v.value / 136e9) if "modmul" in name else "")
PY
KZGB200_LIB=$lib timeout 300 python tools/msm_phases.py 20 21 24 2>&1 | grep msm
KZGB200_LIB=$lib timeout 100 python tools/ntt_once.py 24 2>&1 | tail -2
done 2>&1 | tee gpurun_out/r02_s2c6_karatsuba.log
