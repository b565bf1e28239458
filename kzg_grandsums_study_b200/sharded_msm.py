"""Multi-GPU MSM: one process per GPU, the points and scalars split in contiguous shards.

Rank g owns SRS points and scalars [first_g, first_g + count_g); it reduces its shard to ONE partial point
(extended-Jacobian XYZZ, 128 bytes, in device memory), the partials are exchanged with a single all-gather
(NCCL over NVLink; 128 B per rank, so the collective is pure latency) and every rank sums them and normalises.
Nothing else of the prover shards at these sizes (NTT / scans / whole proofs: replicas only).
"""
import ctypes as C

PARTIAL_WORDS = 16          # 128-byte XYZZ partial as 16 x int64


def shard_range(n, world, rank):
    """contiguous split of n items over `world` ranks; the first n % world ranks take one more"""
    base, extra = divmod(n, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


class ShardedSrsMsm:
    """MSM over a sharded, device-resident SRS.

    `partial_fn(scalars_handle, count, out_tensor)` fills the rank's 128-byte partial; `combine_fn(gathered, world)`
    returns the 64-byte affine result.  The defaults call libkzgb200.so (kzg_srs_msm_partial /
    kzg_g1_partials_combine); tests on CPU (gloo) inject oracle-backed functions to exercise the plumbing.
    """

    def __init__(self, world, rank, device, group=None, curve=None, srs=None, partial_fn=None, combine_fn=None):
        import torch
        self.torch = torch
        self.world, self.rank, self.group = world, rank, group
        self.curve, self.srs = curve, srs
        self.partial = torch.zeros(PARTIAL_WORDS, dtype=torch.int64, device=device)
        self.gathered = torch.zeros(PARTIAL_WORDS * world, dtype=torch.int64, device=device)
        self.partial_fn = partial_fn or self._partial_device
        self.combine_fn = combine_fn or self._combine_device

    def _partial_device(self, scalars, count, out):
        """`scalars`: a device buffer handle (resident shard) or a HOST buffer / tensor (end-to-end: the upload is
        inside, piecewise, hidden behind the pieces' MSMs)"""
        from ._lib import as_ptr
        c = self.curve
        if isinstance(scalars, C.c_void_p) or isinstance(scalars, int):
            c.check(c.lib.kzg_srs_msm_partial(c.ctx, self.srs, 0, scalars, count, as_ptr(out)))
        else:
            c.check(c.lib.kzg_srs_msm_host_partial(c.ctx, self.srs, 0, as_ptr(scalars), count, as_ptr(out)))

    def _combine_device(self, gathered, world):
        from ._lib import as_ptr
        c = self.curve
        out = bytearray(64)
        c.check(c.lib.kzg_g1_partials_combine(c.ctx, as_ptr(gathered), world, as_ptr(out)))
        return bytes(out)

    def _torch_stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def msm(self, scalars, count):
        """scalars: this rank's shard (device buffer handle or host buffer, standard-form LE); returns the 64-byte
        affine sum over ALL ranks' shards (identical on every rank).

        Three steps on two streams: the shard's MSM on the library context's stream, the all-gather on torch's current
        stream (NCCL orders itself against it), the combine on the context's stream again.  The order is made explicit
        with one event each way (kzg_stream_wait_ctx / kzg_ctx_wait_stream) -- a context on a private stream would
        otherwise let the collective read a partial that is still being computed."""
        self.partial_fn(scalars, count, self.partial)
        if self.world == 1:
            return self.combine_fn(self.partial, 1)
        import torch.distributed as dist
        on_device = self.curve is not None and self.partial.is_cuda
        if on_device:
            c = self.curve
            c.check(c.lib.kzg_stream_wait_ctx(c.ctx, self._torch_stream()))   # gather after the partial is complete
        dist.all_gather_into_tensor(self.gathered, self.partial, group=self.group)
        if on_device:
            c.check(c.lib.kzg_ctx_wait_stream(c.ctx, self._torch_stream()))   # combine after the gather has landed
        return self.combine_fn(self.gathered, self.world)
