#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 backend (BASELINE.json metric):
    BN254 G1 MSM Mpts/s at 2^24 points on N GPUs  (value / e2e / roofline)
    + grand-sum prove ms at n = 2^20 through the drop-in prover entry point (extra keys `prove`).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      (CPU arm: the oracle's C restatement on the host cores)

A "step" is one MSM over 2^24 (scalar, SRS point) pairs.  `value` times it with scalars and SRS resident
in HBM; `e2e` goes through the C-ABI call a host binding makes (kzg_g1_msm_affine) with the scalars in
pinned HOST memory: H2D of the scalars and D2H of the 64-byte result are inside the timed region.  With N > 1
the points and scalars are split in N contiguous shards (strong scaling: the total stays 2^24); every rank
reduces its shard to one partial point, the partials (128 B each) are all-gathered over NCCL and summed.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TAU_SEED = 1001
SCALAR_SEED = 6
PROVE_SEED = 4           # config C4
MACS_PER_POINT_WINDOW = 10 * 136   # XYZZ mixed add = 8M + 2S = 10 modmul x 136 limb-MACs


def ncu_counted(key, family):
    """profiles/traffic.json (written by tools/summarize_profiles.py from an ncu capture): DRAM bytes and executed wide
    MACs of the named kernel group -- only while the kernel sources still hash to what was profiled, else None"""
    from kzg_grandsums_study_b200 import build as b
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        entry = json.load(open(path)).get(key)
    except Exception:
        return None
    if not entry:
        return None
    now = b.source_digest(b.MSM_SOURCES if family == "msm" else b.NTT_SOURCES)
    return entry if entry.get("src_sha16") == now else None


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-n", type=int, default=24, help="MSM size (log2), total over all GPUs")
    ap.add_argument("--prove-log-n", type=int, default=20, help="grand-sum prove size (log2); 0 = skip")
    ap.add_argument("--prove-steps", type=int, default=3)
    ap.add_argument("--window", type=int, default=0, help="MSM window bits (0 = library default)")
    ap.add_argument("--table-window", type=int, default=0, help="window bits of the SRS table (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-log-n", type=int, default=0, help="MSM size of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the 2^16..2^24 sweep lines (raw and table MSM)")
    ap.add_argument("--prove-log-n2", type=int, default=22, help="second prove size reported under prove.also (0 = skip)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for (t, r) in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1 + 0.2)] or \
               [r for (_, r) in self.rows]
        sm, mx, reasons, power = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                power.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------------------------------------------
# CPU arm (oracle's C restatement; test infrastructure used as the baseline only)
# ------------------------------------------------------------------------------------------------------
def cpu_msm_sample(log_n, threads=None):
    """time the oracle's multi-threaded Pippenger on a 2^log_n sample; returns (Mpts/s, seconds, threads)"""
    from oracle.c import binding as ob
    return ob.bench_msm(log_n, TAU_SEED, SCALAR_SEED, threads)


def run_reference(args):
    """CPU arm: the oracle's C restatement of the reference's multiExpAffine (ffjavascript's unsigned-window Pippenger
    with its pTSizes table) on ALL host cores.  The real reference (ffjavascript WASM under Node) cannot run on this
    image; bench/ref_node/time_msm.js is the same measurement for a machine that has Node.
    Same configuration as the GPU arm whenever the box finishes it in a few minutes: the FULL 2^log_n MSM per step
    (~5-10 s on 16-32 cores); otherwise a smaller sample, named in `config`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import ctypes as C2
    from oracle.c import binding as ob
    from oracle.py import inputs
    from kzg_grandsums_study_b200 import synthetic
    cores = os.cpu_count() or 1
    budget_s = 200.0
    sample_log = args.cpu_sample_log_n or args.log_n
    lib = ob.lib()
    tau = inputs.tau_from_seed(TAU_SEED)
    # a quick 2^18 probe sizes the run: the full size only if K + W steps (and the SRS generation) fit the budget
    mpts_probe, _, _ = ob.bench_msm(18, TAU_SEED, SCALAR_SEED, cores)
    steps_total = args.steps + (1 if args.warmup > 0 else 0)
    if not args.cpu_sample_log_n:
        while sample_log > 18 and ((1 << sample_log) / (mpts_probe * 1e6)) * (steps_total + 1.5) > budget_s:
            sample_log -= 2
    n = 1 << sample_log
    bases = C2.create_string_buffer(64 * n)
    lib.ko_srs_generate(ob._buf(int(tau).to_bytes(32, "little")), 0, n, bases, cores)
    scal = ob._buf(synthetic.random_fr_std(SCALAR_SEED, 1 << args.log_n)[:n].tobytes())
    out = C2.create_string_buffer(64)
    times = []
    for i in range(steps_total):
        t0 = time.perf_counter()
        lib.ko_g1_msm(bases, scal, n, out, cores)
        dt = time.perf_counter() - t0
        if i >= steps_total - args.steps:
            times.append(dt)
    t = sum(times) / len(times)
    value = n / t / 1e6
    same = sample_log == args.log_n
    line = {
        "impl": "reference", "metric": "bn254_g1_msm_throughput", "value": value, "unit": "Mpts/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32x8 (BN254 Fq/Fr Montgomery)", "data": "synthetic",
        "config": {"workload": "BN254 G1 MSM 2^%d points (BASELINE config 5), CPU arm on %s" % (
                       args.log_n, "the same 2^%d points" % sample_log if same else "a 2^%d-point sample" % sample_log),
                   "same_config": same, "cpu_sample_log_n": sample_log, "host_cores": cores,
                   "reference_kind": "C/OpenMP port of ffjavascript's Pippenger (oracle/c); the WASM original needs Node "
                                     "(bench/ref_node/time_msm.js)",
                   "result_affine_hex": out.raw.hex()},
        "cpu_baseline": {"value": value, "unit": "Mpts/s", "cores": cores, "kind": "port",
                         "sample": "2^%d of the 2^%d (SRS point, scalar) pairs per step, %d threads, oracle/c Pippenger "
                                   "(ffjavascript window table)" % (sample_log, args.log_n, cores)},
        "e2e": {"value": value, "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from kzg_grandsums_study_b200 import _lib, synthetic
    from kzg_grandsums_study_b200._lib import as_ptr
    from kzg_grandsums_study_b200 import curve as curve_mod
    from kzg_grandsums_study_b200.curve import Curve

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torchrun --nproc-per-node %d" % (args.gpus, world, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    # a non-default torch stream made current for the whole run: the library launches on it (a NULL stream
    # pointer would make kzg_ctx_create open a private stream that torch's events do not see)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    curve = Curve(local_rank, stream.cuda_stream)     # kernels go on torch's current stream: torch events see them
    curve_mod._CURVES[local_rank] = curve             # the drop-in provers pick this curve up (getCurveFromName cache)
    lib, ctx = curve.lib, curve.ctx
    if args.window:
        curve.check(lib.kzg_msm_set_window(ctx, args.window))

    N = 1 << args.log_n
    shard = N // world
    first = rank * shard
    tau = synthetic.tau_from_seed(TAU_SEED)

    # ---- resident inputs: this rank's SRS shard and scalar shard ----
    srs = C.c_void_p()
    curve.check(lib.kzg_srs_generate_range(ctx, as_ptr(tau.to_bytes(32, "little")), first, shard, C.byref(srs)))
    srs_precompute_ms = None
    if not args.window:
        torch.cuda.synchronize()
        tp = time.perf_counter()
        curve.check(lib.kzg_srs_precompute(ctx, srs, args.table_window))   # one-off, with the SRS, before any timing
        torch.cuda.synchronize()
        srs_precompute_ms = (time.perf_counter() - tp) * 1e3
    scal_all = synthetic.random_fr_std(SCALAR_SEED, N)             # standard-form LE scalars, (N, 4) u64
    scal_host = torch.from_numpy(scal_all[first:first + shard].view(np.int64).copy()).pin_memory()
    del scal_all
    scal_dev = curve.alloc(shard)
    curve.check(lib.kzg_buf_upload(ctx, scal_dev.handle, 0, as_ptr(scal_host), shard))
    from kzg_grandsums_study_b200.sharded_msm import ShardedSrsMsm
    sharded = ShardedSrsMsm(world, rank, dev, curve=curve, srs=srs)
    out_affine = bytearray(64)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        if world == 1:
            curve.check(lib.kzg_srs_msm(ctx, srs, 0, scal_dev.handle, shard, as_ptr(out_affine)))
        else:
            out_affine[:] = sharded.msm(scal_dev.handle, shard)

    def step_e2e():
        # the call a host binding makes: scalars in (pinned) host memory, bases = the resident SRS shard
        if world == 1:
            curve.check(lib.kzg_srs_msm_host(ctx, srs, 0, as_ptr(scal_host), shard, as_ptr(out_affine)))
        else:
            out_affine[:] = sharded.msm(scal_host, shard)   # kzg_srs_msm_host_partial: the upload is inside, piecewise

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            flush.fill_(1)
            fn()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        t0 = time.time()
        for a, b in evs:
            flush.fill_(1)              # L2 flush between timed iterations (outside the event pair)
            a.record()
            fn()
            b.record()
        barrier()
        t1 = time.time()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), t0, t1

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- value: resident inputs ----
    for _ in range(args.warmup):
        step_resident()
    curve.check(lib.kzg_ctx_kernel_time(ctx, 0, 1, None, None))       # clear + enable per-kernel timing
    launches0 = curve.launch_count()
    total_ms, t0, t1 = timed(step_resident, args.steps, 0)
    launches = curve.launch_count() - launches0
    acc_ms = C.c_double()
    acc_launches = C.c_uint64()
    aff_ms = C.c_double()
    aff_phases = C.c_uint64()
    curve.check(lib.kzg_ctx_kernel_time(ctx, 5, 0, C.byref(aff_ms), C.byref(aff_phases)))      # batched-affine rounds
    curve.check(lib.kzg_ctx_kernel_time(ctx, 0, -1, C.byref(acc_ms), C.byref(acc_launches)))   # XYZZ walk; stop timing
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = N / (ms_per_step * 1e-3) / 1e6
    result_resident = bytes(out_affine)

    # ---- e2e: scalars from pinned host memory through the C-ABI entry point ----
    e2e_total, _, _ = timed(step_e2e, args.steps, 1)
    e2e_ms = e2e_total / args.steps
    if bytes(out_affine) != result_resident:
        raise SystemExit("bench.py: e2e result differs from the resident-input result")

    # ---- known answer: (sum s_i tau^i) G1 -- per-shard sums by the library's own Horner kernel, added on the host ----
    k_shard = shard_known_scalar(curve, scal_dev, tau, first)
    if world > 1:
        ks = torch.tensor([(k_shard >> (62 * j)) & ((1 << 62) - 1) for j in range(5)], dtype=torch.int64, device=dev)
        allk = torch.zeros(5 * world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allk, ks)
        allk = allk.cpu().tolist()
        k_total = sum(sum(allk[5 * g + j] << (62 * j) for j in range(5)) for g in range(world)) % curve.r
    else:
        k_total = k_shard
    check = known_answer_matches(curve, k_total, result_resident)
    if not check:
        raise SystemExit("bench.py: MSM result does not match the closed form (sum s_i tau^i) G1")

    # ---- roofline of the dominant kernel (bucket accumulation), measured live ----
    imad = C.c_double()
    modmul = C.c_double()
    curve.check(lib.kzg_bench_imad_peak(ctx, 200, C.byref(imad)))
    curve.check(lib.kzg_bench_modmul_peak(ctx, 200, C.byref(modmul)))
    geom = msm_geometry(lib, ctx, srs, shard)
    # The dominant work is the BUCKET ACCUMULATION: `affine_rounds` batched-affine rounds (msm_aff_forward /
    # fq_batch_inverse / msm_aff_backward, tag 5) that halve the sorted list each time, then the XYZZ walk
    # msm_accumulate_kernel (tag 0) over what is left.  One "launch" below = one accumulation = one MSM.
    phase_ms = (acc_ms.value + aff_ms.value) / max(1, acc_launches.value)
    walk_ms = acc_ms.value / max(1, acc_launches.value)
    # peak: wide MACs per second.  The raw IMAD.WIDE chain microbenchmark and the library's own product chain (136 counted
    # MACs per product, 128 of them wide) are two measurements of the same pipe; the larger one is the denominator.
    peak = max(imad.value, modmul.value * 128.0 / 136.0)
    # ALGORITHMIC work (the progress metric): 21 760 limb-MACs per point = 16 windows x one XYZZ mixed add.
    macs_per_launch = float(shard) * 16 * MACS_PER_POINT_WINDOW
    algorithmic = macs_per_launch / (phase_ms * 1e-3) if phase_ms > 0 else 0.0
    # EXECUTED work: wide MACs the kernels really issue.  Counted by ncu (fmaheavy-pipe thread instructions of these
    # kernels at this size, profiles/traffic.json, valid while the kernel sources hash the same); else modelled: an entry
    # absorbed by an affine round 778 wide MACs (forward product 128, backward 4 x 128 + one squaring of 100, ~0.3 products
    # for its share of the recursive inversion), an entry left to the walk 1160 (6 products, 2 squarings, R (Q - X3) -
    # Y1 PPP as two products under one reduction = 192).
    counted = ncu_counted("msm_accumulation_2_%d" % args.log_n, "msm") if world == 1 else None
    entries = float(shard) * geom["windows"]
    left = entries / (1 << geom["affine_rounds"])
    modelled_macs = left * 1160.0 + (entries - left) * 778.0
    executed_macs = counted["fmaheavy_thread_instructions"] if counted else modelled_macs
    executed = executed_macs / (phase_ms * 1e-3) if phase_ms > 0 else 0.0
    peaks, peak_kind = measured_peaks()
    affine = geom["affine_rounds"] > 0
    roofline = {
        "bound": "imad",
        "kernel": ("bucket accumulation = %d batched-affine rounds (msm_aff_forward_kernel, batch_*_kernel<FqP>, "
                   "msm_aff_backward_kernel) + msm_accumulate_kernel" % geom["affine_rounds"]) if affine else "msm_accumulate_kernel",
        # frac = EXECUTED wide MACs / kernel time / measured pipe peak: a utilisation figure (<= 1)
        "achieved": executed / 1e12, "peak": peak / 1e12, "unit": "Tmac/s", "frac": executed / peak if peak else None,
        "executed_macs_source": "ncu (profiles/traffic.json)" if counted else "model (778 / 1160 wide MACs per entry)",
        "traffic": counted["dram_bytes"] if counted else None,
        "algorithmic_bytes": 64.0 * entries,
        # the SURVEY 8d unit (21 760 MACs per point whatever the kernel really does): > 1 means algorithmic savings
        # (window table: 12-13 digits instead of 16; affine additions: ~6 products instead of 10), not a saturated pipe
        "algorithmic_tmacs": algorithmic / 1e12, "algorithmic_frac": algorithmic / peak if peak else None,
        "peak_source": "max(kzg_bench_imad_peak, kzg_bench_modmul_peak x 128/136): IMAD.WIDE.U32 chains timed live on this GPU "
                       "(MEASURED_PEAKS.json has no integer peak)",
        "imad_peak_tmacs": imad.value / 1e12, "modmul_peak_tmacs": modmul.value / 1e12,
        "algorithmic_macs_per_launch": macs_per_launch, "executed_macs_per_launch": executed_macs,
        "modelled_macs_per_launch": modelled_macs,
        "kernel_ms_per_launch": phase_ms,
        "affine_rounds_ms_per_launch": aff_ms.value / max(1, acc_launches.value),
        "xyzz_walk_ms_per_launch": walk_ms,
        "kernel_share_of_step": phase_ms / ms_per_step if ms_per_step else None,
        "windows": geom["windows"], "window_bits": geom["c"], "affine_rounds": geom["affine_rounds"],
        "hbm_peak_gbs": peaks.get("hbm_gbs"), "hbm_peak_kind": peak_kind,
    }

    # ---- the HBM-side kernels: one NTT of the size the n = 2^22 prover needs (2^24), timed live (rank 0) ----
    roofline_hbm = None
    if rank == 0 and world == 1:
        roofline_hbm = bench_ntt(curve, 24, peaks.get("hbm_gbs"), peak_kind, peak)

    # ---- what the reference's own calls cost at every size of BASELINE config 5 (rank 0, N = 1) ----
    sweep = None
    if rank == 0 and world == 1 and not args.no_sweep:
        sweep = bench_sweep(curve, tau, srs, scal_dev, args.log_n)

    # ---- grand-sum prove at n = 2^prove_log_n through the drop-in entry point (rank 0 only) ----
    prove = None
    if args.prove_log_n and rank == 0:
        prove = bench_prove(curve, args.prove_log_n, args.prove_steps, tau)
        if args.prove_log_n2 and args.prove_log_n2 != args.prove_log_n and world == 1:
            prove["also"] = bench_prove(curve, args.prove_log_n2, 2, tau)   # BASELINE config 4's second size
    if world > 1:
        dist.barrier()

    # ---- CPU baseline beside it (rank 0, N = 1) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sample_log = args.cpu_sample_log_n or min(args.log_n, 20)
            mpts, secs, thr = cpu_msm_sample(sample_log)
            cpu = {"value": mpts, "unit": "Mpts/s", "cores": thr, "kind": "port",
                   "sample": "2^%d-point MSM (same SRS and scalar prefix), oracle/c Pippenger with the ffjavascript window "
                             "table on %d threads, %.2f s" % (sample_log, thr, secs)}
        except Exception as e:  # the baseline is a reported side figure; its absence must not hide the GPU numbers
            cpu = {"value": None, "unit": "Mpts/s", "cores": 0, "kind": "port", "sample": "unavailable: %s" % e}
        if prove is not None:
            try:
                from oracle.c import binding as ob
                cpu_log = min(args.prove_log_n, 20)
                secs, thr, digest = ob.bench_prove(cpu_log, TAU_SEED, PROVE_SEED)
                prove["cpu_baseline"] = {"value": secs * 1e3, "unit": "ms", "cores": thr, "kind": "port", "n": 1 << cpu_log,
                                         "sample": "one grand-sum proof at n=2^%d by oracle/c (the reference's algorithm: NTT "
                                                   "multiplications at 2n/4n, divZh, ffjavascript-style Pippenger) on %d threads"
                                                   % (cpu_log, thr)}
                if cpu_log == args.prove_log_n:
                    prove["byte_identical_to_cpu_oracle"] = digest == prove["proof_sha256"]
            except Exception as e:
                prove["cpu_baseline"] = {"value": None, "unit": "ms", "cores": 0, "kind": "port", "sample": "unavailable: %s" % e}

    if rank == 0:
        line = {
            "metric": "bn254_g1_msm_throughput", "value": value, "unit": "Mpts/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32x8 (BN254 Fq/Fr Montgomery)", "data": "synthetic",
            "config": {"workload": "BN254 G1 MSM, 2^%d points total (BASELINE config 5), SRS and scalars resident" % args.log_n,
                       "points_per_gpu": shard, "parallelism": "msm-shard%d" % world,
                       "flavour": "fixed-base window table built once per resident SRS (kzg_srs_precompute), excluded "
                                  "from the step like the SRS load; `sweep` has the table-less numbers",
                       "srs_precompute_ms": srs_precompute_ms,
                       "table_bytes": 64 * shard * geom["windows"] if not args.window else 0, "host_cores": os.cpu_count(),
                       "l2": "256 MiB buffer rewritten between timed iterations; inputs (1.5 GiB) exceed L2",
                       "known_answer_ok": check, "result_affine_hex": result_resident.hex()},
            "e2e": {"value": N / (e2e_ms * 1e-3) / 1e6, "unit": "Mpts/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 32 * shard, "d2h_bytes_per_step": 64},
            "gpu_launches": int(launches),
            "clocks": clocks, "roofline": roofline, "roofline_ntt": roofline_hbm, "cpu_baseline": cpu, "prove": prove,
            "sweep": sweep,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def msm_geometry(lib, ctx, srs, n):
    c = C.c_uint32()
    w = C.c_uint32()
    r = C.c_uint32()
    lib.kzg_msm_plan(ctx, srs, n, 0, C.byref(c), C.byref(w), C.byref(r))
    return {"c": c.value, "windows": w.value, "affine_rounds": r.value}


def shard_known_scalar(curve, scal_dev, tau, first):
    """sum_i s_i tau^(first + i) mod r for this rank's shard, by the device Horner evaluation.  The standard-form
    scalars are re-read as Montgomery residues (value s_i / 2^256) and the result comes back as Montgomery bytes,
    i.e. the raw little-endian integer is exactly sum s_i tau^i."""
    from kzg_grandsums_study_b200._lib import as_ptr
    R = curve.r
    out = bytearray(32)
    tau_m = (tau << 256) % R
    curve.check(curve.lib.kzg_poly_evaluate(curve.ctx, scal_dev.handle, as_ptr(tau_m.to_bytes(32, "little")), as_ptr(out)))
    return int.from_bytes(out, "little") % R * pow(tau, first, R) % R


def known_answer_matches(curve, k, got_affine):
    """k * G1 by a 1-point MSM on the generator == got_affine"""
    from kzg_grandsums_study_b200._lib import as_ptr
    gen = ((1 << 256) % curve.q).to_bytes(32, "little") + ((2 << 256) % curve.q).to_bytes(32, "little")
    aff = bytearray(64)
    curve.check(curve.lib.kzg_g1_msm_affine(curve.ctx, as_ptr(gen), as_ptr(k.to_bytes(32, "little")), 1, 0, as_ptr(aff), None))
    return bytes(aff) == bytes(got_affine)


def bench_ntt(curve, log_n, hbm_peak, peak_kind, imad_peak):
    """Fr NTT of 2^log_n elements, CUDA events on the launching stream.  Bound by the integer pipe, not by HBM: a 254-bit
    butterfly is one 136-MAC product per two elements and stage, against 64 bytes per element and pass."""
    import torch
    from kzg_grandsums_study_b200 import synthetic
    n = 1 << log_n
    x = curve.to_device(synthetic.random_fr_std(77, n).tobytes())
    y = curve.alloc(n)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    best = {0: None, 1: None}
    for inverse in (0, 1):
        for i in range(6):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            curve.check(curve.lib.kzg_fr_ntt(curve.ctx, x.handle, y.handle, inverse))
            b.record()
            torch.cuda.synchronize()
            if i >= 2:
                t = a.elapsed_time(b)
                best[inverse] = t if best[inverse] is None else min(best[inverse], t)
    ms = best[0]
    counted = ncu_counted("ntt_2_%d" % log_n, "ntt")
    # model: per pass ~3 products per element in the butterflies (trivial twiddles skipped by whole warps), 1 at every pass
    # boundary (direct twiddle tables; N^-1 of an inverse transform rides on the last one)
    passes = (log_n + 7) // 8
    modelled = n * (3.0 * passes + (passes - 1)) * 128
    executed = counted["fmaheavy_thread_instructions"] if counted else modelled
    rate = executed / (ms * 1e-3)
    return {"bound": "imad", "kernel": "ntt_strided_pass_kernel x%d + ntt_last_pass_kernel (one 2^%d transform)" % (passes - 1, log_n),
            "achieved": rate / 1e12, "peak": imad_peak / 1e12, "unit": "Tmac/s", "frac": rate / imad_peak if imad_peak else None,
            "executed_macs_source": "ncu (profiles/traffic.json)" if counted else "model",
            "modmul_per_element": executed / 128.0 / n, "ms": ms, "inverse_ms": best[1],
            "traffic": counted["dram_bytes"] if counted else None, "algorithmic_bytes": 64.0 * n,
            "hbm_gbs_algorithmic": 64.0 * n / (ms * 1e-3) / 1e9, "hbm_peak_gbs": hbm_peak, "hbm_peak_kind": peak_kind,
            "hbm_frac_algorithmic": 64.0 * n / (ms * 1e-3) / 1e9 / hbm_peak if hbm_peak else None,
            "note": "%d shared-memory passes (TMA bulk row loads, radix-4 register butterflies); SURVEY 8d files the NTT "
                    "under HBM, the counters say integer pipe (fmaheavy ~80 %%)" % passes}


def bench_sweep(curve, tau, srs_full, scal_full, log_n_max):
    """BASELINE config 5 on one GPU, device timing (CUDA events), inputs resident:
       raw   = kzg_g1_msm_affine on caller-supplied bases -- what G1.multiExpAffine (polynomial.js:1112) binds to: no table
       table = kzg_srs_msm over an SRS of that size with its window table (what commit() uses)"""
    import torch
    from kzg_grandsums_study_b200._lib import as_ptr, KZG_BASES_ON_DEVICE, KZG_SCALARS_ON_DEVICE
    lib, ctx = curve.lib, curve.ctx
    out = bytearray(64)
    bases_ptr = lib.kzg_srs_device_ptr(srs_full)
    scal_ptr = lib.kzg_buf_device_ptr(scal_full.handle)

    def timed(fn, reps=3):
        fn()
        best = None
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            t = a.elapsed_time(b)
            best = t if best is None else min(best, t)
        return best

    rows = []
    for lg in (16, 18, 20, 22, 24):
        if lg > log_n_max:
            break
        n = 1 << lg
        row = {"log_n": lg}
        ms = timed(lambda: curve.check(lib.kzg_g1_msm_affine(ctx, bases_ptr, scal_ptr, n, KZG_BASES_ON_DEVICE | KZG_SCALARS_ON_DEVICE,
                                                             as_ptr(out), None)))
        row["raw_ms"] = ms
        row["raw_mpts_s"] = n / ms / 1e3
        raw_result = bytes(out)
        if lg <= 22:
            s = C.c_void_p()
            curve.check(lib.kzg_srs_generate(ctx, as_ptr(tau.to_bytes(32, "little")), n, C.byref(s)))
            t0 = time.perf_counter()
            curve.check(lib.kzg_srs_precompute(ctx, s, 0))
            torch.cuda.synchronize()
            row["table_precompute_ms"] = (time.perf_counter() - t0) * 1e3
            ms = timed(lambda: curve.check(lib.kzg_srs_msm(ctx, s, 0, scal_full.handle, n, as_ptr(out))))
            row["table_ms"] = ms
            row["table_mpts_s"] = n / ms / 1e3
            row["same_point"] = bytes(out) == raw_result
            lib.kzg_srs_free(ctx, s)
        rows.append(row)
    return {"what": "BN254 G1 MSM on 1 GPU, resident inputs, best of 3 (CUDA events): raw = kzg_g1_msm_affine on caller bases "
                    "(no table), table = kzg_srs_msm with the SRS window table", "rows": rows}


def bench_prove(curve, log_n, steps, tau):
    """grand-sum multiset-equality prove (BASELINE config 4) through mset_eq_kzg_grandsum_prover: host columns in,
    proof out; SRS resident (loaded once from the synthetic .ptau, like the reference's PTau buffer)"""
    import tempfile

    import numpy as np
    import torch

    from kzg_grandsums_study_b200 import synthetic
    from kzg_grandsums_study_b200._lib import as_ptr
    from kzg_grandsums_study_b200.grandsum import mset_eq_kzg_grandsum_prover
    from kzg_grandsums_study_b200.polynomial import Evaluations
    lib, ctx = curve.lib, curve.ctx
    n = 1 << log_n
    f = synthetic.random_fr_std(PROVE_SEED, n)
    perm = synthetic.permutation(PROVE_SEED, n)
    t = f[perm]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "bench_%d.ptau" % log_n)
        srs = C.c_void_p()
        curve.check(lib.kzg_srs_generate(ctx, as_ptr(tau.to_bytes(32, "little")), 2 * n, C.byref(srs)))
        zero128 = bytes(128)        # the prover never reads section 3
        curve.check(lib.kzg_srs_write_ptau(ctx, srs, log_n, as_ptr(zero128), as_ptr(zero128), path.encode()))
        lib.kzg_srs_free(ctx, srs)
        # the caller's columns in PINNED host memory (the H2D copies are inside the timed region)
        fb = torch.from_numpy(f.view(np.uint8).reshape(-1).copy()).pin_memory()
        tb = torch.from_numpy(np.ascontiguousarray(t).view(np.uint8).reshape(-1).copy()).pin_memory()
        times = []
        launches = 0
        proof = None
        for i in range(steps + 1):
            torch.cuda.synchronize()
            l0 = curve.launch_count()
            t0 = time.perf_counter()
            proof = mset_eq_kzg_grandsum_prover(path, Evaluations(fb, curve), Evaluations(tb, curve))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if i > 0:               # the first call also loads the .ptau into HBM
                times.append(dt)
                launches = curve.launch_count() - l0
        times.sort()
        import hashlib
        digest = hashlib.sha256(b"".join(proof["commitments"].values()) + b"".join(proof["evaluations"].values())).hexdigest()
        return {"metric": "grandsum_prove_ms", "n": n, "median_ms": times[len(times) // 2] * 1e3, "min_ms": times[0] * 1e3,
                "steps": steps, "gpu_launches": int(launches), "h2d_bytes": 64 * n, "proof_sha256": digest,
                "timing": "host wall clock around the drop-in prover call (columns in pinned host memory, SRS resident)"}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
