"""CPU, world_size 2, gloo: the host logic of the multi-GPU MSM (shard ranges, one all-gather of 128-byte partials,
combine on every rank).  The per-shard arithmetic is injected from the oracle -- on the GPU box the same class runs
with its default libkzgb200.so hooks (bench.py --gpus N)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range():
    from kzg_grandsums_study_b200.sharded_msm import shard_range
    for n in (0, 1, 7, 16, 1 << 24, (1 << 24) + 3):
        for world in (1, 2, 3, 4, 8):
            pos = 0
            for r in range(world):
                first, cnt = shard_range(n, world, r)
                assert first == pos and cnt in (n // world, n // world + 1)
                pos += cnt
            assert pos == n


def _worker(rank, world, port, n, seed, tau, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from kzg_grandsums_study_b200.sharded_msm import ShardedSrsMsm, shard_range
        from oracle.c import binding as oc
        from oracle.py import bn254 as bn, inputs
        first, cnt = shard_range(n, world, rank)
        srs = oc.srs_generate(tau, cnt, first=first, threads=2)          # this rank's SRS shard
        scalars = inputs.random_column(seed, n)[first:first + cnt]         # this rank's scalar shard
        one = bn.fq_to_mont_bytes(1)

        def partial_fn(handle, count, out):
            aff = oc.msm(srs, bn.fr_vec_to_std_bytes(handle), threads=2)   # "handle" is the scalar list on CPU
            xyzz = aff + (bytes(64) if aff == bytes(64) else one + one)    # affine -> XYZZ (zz = zzz = 1; 0 = infinity)
            out.copy_(torch.frombuffer(bytearray(xyzz), dtype=torch.int64))

        def combine_fn(gathered, w):
            raw = gathered.numpy().tobytes()
            acc = None
            for g in range(w):
                part = raw[128 * g:128 * g + 128]
                if part[64:96] != bytes(32):
                    acc = bn.g1_add(acc, bn.g1_from_bytes(part[:64])) if acc is not None else bn.g1_from_bytes(part[:64])
            return bn.g1_to_bytes(acc)

        m = ShardedSrsMsm(world, rank, "cpu", partial_fn=partial_fn, combine_fn=combine_fn)
        got = m.msm(scalars, cnt)
        q.put((rank, got))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_msm():
    from oracle.py import bn254 as bn, inputs
    n, seed, tau = 301, 17, 123456789
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, seed, tau, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    scalars = inputs.random_column(seed, n)
    expect = sum(x * pow(tau, i, bn.R) for i, x in enumerate(scalars)) % bn.R
    want = bn.g1_to_bytes(bn.g1_mul_gen(expect))
    assert results[0] == want and results[1] == want
