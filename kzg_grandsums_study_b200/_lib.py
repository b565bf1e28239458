"""ctypes binding of libkzgb200.so -- exactly the symbols of include/kzgb200.h.

This is the stand-in for the N-API addon on images without Node (INTEGRATION.md shows the addon stub):
plain pointers and sizes only.  There is no CPU fallback: if the library is missing the import fails, and
without a CUDA device `kzg_ctx_create` fails.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KZGB200_LIB") or os.path.join(HERE, "libkzgb200.so")  # (override: A/B builds of experiments)

u8p = C.POINTER(C.c_uint8)
vp = C.c_void_p
u32 = C.c_uint32
u64 = C.c_uint64
i32 = C.c_int

# name -> (restype, argtypes); mirrors include/kzgb200.h one to one
SIGNATURES = {
    "kzg_ctx_create": (i32, [i32, vp, C.POINTER(vp)]),
    "kzg_ctx_destroy": (i32, [vp]),
    "kzg_ctx_sync": (i32, [vp]),
    "kzg_ctx_wait_stream": (i32, [vp, vp]),
    "kzg_stream_wait_ctx": (i32, [vp, vp]),
    "kzg_ctx_set_option": (i32, [vp, C.c_char_p, C.c_int64]),
    "kzg_last_error": (C.c_char_p, [vp]),
    "kzg_ctx_launch_count": (u64, [vp]),
    "kzg_selftest": (i32, [vp, u32]),
    "kzg_bench_imad_peak": (i32, [vp, u32, C.POINTER(C.c_double)]),
    "kzg_bench_modmul_peak": (i32, [vp, u32, C.POINTER(C.c_double)]),
    "kzg_ctx_kernel_time": (i32, [vp, u32, i32, C.POINTER(C.c_double), C.POINTER(u64)]),
    "kzg_srs_load_ptau": (i32, [vp, C.c_char_p, u64, C.POINTER(vp), C.POINTER(u32)]),
    "kzg_srs_load_ptau_range": (i32, [vp, C.c_char_p, u64, u64, C.POINTER(vp), C.POINTER(u32)]),
    "kzg_ptau_read_header": (i32, [vp, C.c_char_p, C.POINTER(u32), C.POINTER(u32)]),
    "kzg_ptau_read_tau_g2": (i32, [vp, C.c_char_p, vp]),
    "kzg_srs_from_host": (i32, [vp, vp, u64, C.POINTER(vp)]),
    "kzg_srs_generate": (i32, [vp, vp, u64, C.POINTER(vp)]),
    "kzg_srs_generate_range": (i32, [vp, vp, u64, u64, C.POINTER(vp)]),
    "kzg_srs_lagrange": (i32, [vp, vp, u32, C.POINTER(vp)]),
    "kzg_srs_write_ptau": (i32, [vp, vp, u32, vp, vp, C.c_char_p]),
    "kzg_srs_download": (i32, [vp, vp, u64, u64, vp]),
    "kzg_srs_len": (u64, [vp]),
    "kzg_srs_device_ptr": (vp, [vp]),
    "kzg_srs_free": (i32, [vp, vp]),
    "kzg_buf_alloc": (i32, [vp, u64, C.POINTER(vp)]),
    "kzg_buf_free": (i32, [vp, vp]),
    "kzg_buf_len": (u64, [vp]),
    "kzg_buf_device_ptr": (vp, [vp]),
    "kzg_buf_upload": (i32, [vp, vp, u64, vp, u64]),
    "kzg_buf_download": (i32, [vp, vp, u64, vp, u64]),
    "kzg_buf_copy": (i32, [vp, vp, u64, vp, u64, u64]),
    "kzg_buf_fill": (i32, [vp, vp, u64, u64, vp]),
    "kzg_buf_all_equal": (i32, [vp, vp, vp, C.POINTER(i32)]),
    "kzg_fr_to_mont": (i32, [vp, vp, vp]),
    "kzg_fr_from_mont": (i32, [vp, vp, vp]),
    "kzg_fr_ntt": (i32, [vp, vp, vp, i32]),
    "kzg_fr_extend_ntt": (i32, [vp, vp, u32, C.POINTER(vp)]),
    "kzg_fr_batch_inverse": (i32, [vp, vp, vp]),
    "kzg_poly_add": (i32, [vp, vp, vp, C.POINTER(vp)]),
    "kzg_poly_sub": (i32, [vp, vp, vp, C.POINTER(vp)]),
    "kzg_poly_mul_scalar": (i32, [vp, vp, vp]),
    "kzg_poly_add_scalar": (i32, [vp, vp, vp]),
    "kzg_poly_sub_scalar": (i32, [vp, vp, vp]),
    "kzg_poly_degree": (i32, [vp, vp, C.POINTER(u64)]),
    "kzg_poly_evaluate": (i32, [vp, vp, vp, vp]),
    "kzg_poly_multiply": (i32, [vp, vp, vp, C.POINTER(vp)]),
    "kzg_poly_shift_omega": (i32, [vp, vp, C.POINTER(vp)]),
    "kzg_poly_div_zh": (i32, [vp, vp, u64, C.POINTER(vp)]),
    "kzg_poly_div_x_sub_value": (i32, [vp, vp, vp, C.POINTER(vp)]),
    "kzg_poly_lagrange1": (i32, [vp, u32, C.POINTER(vp)]),
    "kzg_grandsum_build": (i32, [vp, vp, vp, vp, vp, vp, C.POINTER(vp)]),
    "kzg_grandproduct_build": (i32, [vp, vp, vp, vp, vp, vp, C.POINTER(vp)]),
    "kzg_commit": (i32, [vp, vp, vp, vp]),
    "kzg_g1_msm_affine": (i32, [vp, vp, vp, u64, u32, vp, vp]),
    "kzg_srs_msm": (i32, [vp, vp, u64, vp, u64, vp]),
    "kzg_srs_msm_partial": (i32, [vp, vp, u64, vp, u64, vp]),
    "kzg_g1_partials_combine": (i32, [vp, vp, u32, vp]),
    "kzg_srs_msm_host": (i32, [vp, vp, u64, vp, u64, vp]),
    "kzg_srs_msm_host_partial": (i32, [vp, vp, u64, vp, u64, vp]),
    "kzg_commit_many": (i32, [vp, vp, C.POINTER(vp), u32, vp]),
    "kzg_srs_precompute": (i32, [vp, vp, u32]),
    "kzg_msm_geometry": (i32, [vp, vp, u64, i32, C.POINTER(u32), C.POINTER(u32)]),
    "kzg_msm_plan": (i32, [vp, vp, u64, i32, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]),
    "kzg_msm_set_window": (i32, [vp, u32]),
    "kzg_mgpu_create": (i32, [C.POINTER(i32), u32, C.POINTER(vp)]),
    "kzg_mgpu_destroy": (i32, [vp]),
    "kzg_mgpu_device_count": (u32, [vp]),
    "kzg_mgpu_ctx": (vp, [vp, u32]),
    "kzg_mgpu_last_error": (C.c_char_p, [vp]),
    "kzg_mgpu_srs_len": (u64, [vp]),
    "kzg_mgpu_shard": (i32, [vp, u32, C.POINTER(u64), C.POINTER(u64)]),
    "kzg_mgpu_srs_generate": (i32, [vp, vp, u64]),
    "kzg_mgpu_srs_from_host": (i32, [vp, vp, u64]),
    "kzg_mgpu_srs_load_ptau": (i32, [vp, C.c_char_p, u64]),
    "kzg_mgpu_srs_msm_host": (i32, [vp, vp, u64, vp]),
    "kzg_mgpu_scalars_upload": (i32, [vp, vp, u64]),
    "kzg_mgpu_srs_msm": (i32, [vp, vp]),
    "kzg_host_register": (i32, [vp, u64]),
    "kzg_host_unregister": (i32, [vp]),
    "kzg_prover_create": (i32, [vp, vp, i32, u32, u32, i32, C.POINTER(vp)]),
    "kzg_prover_destroy": (i32, [vp]),
    "kzg_prover_round1": (i32, [vp, C.POINTER(vp), C.POINTER(vp), vp, vp, vp]),
    "kzg_prover_round2": (i32, [vp, vp, vp, vp]),
    "kzg_prover_round3": (i32, [vp, vp, vp]),
    "kzg_prover_round4": (i32, [vp, vp, vp]),
    "kzg_prover_round5": (i32, [vp, vp, vp]),
    "kzg_prover_take_evals": (i32, [vp, u32, i32, C.POINTER(vp)]),
    "kzg_prover_last_error": (C.c_char_p, [vp]),
    "kzg_prover_n_evals": (u32, [vp]),
    "kzg_prover_n_round1_commitments": (u32, [vp]),
    "kzg_keccak256": (None, [vp, C.c_size_t, vp]),
    "kzg_g1_to_rpr_uncompressed": (None, [vp, vp]),
    "kzg_fr_to_rpr_be": (None, [vp, vp]),
    "kzg_fr_from_hash_be": (None, [vp, vp]),
}

KZG_GRANDSUM, KZG_GRANDPRODUCT = 0, 1
KZG_BASES_ON_DEVICE, KZG_SCALARS_ON_DEVICE = 1, 2
KZG_ERR_PROTOCOL = -5

_lib = None


def load():
    """Load libkzgb200.so (built in-tree by kzg_grandsums_study_b200.build) and type every symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libkzgb200.so is not built (%s). Run `python -m kzg_grandsums_study_b200.build`; "
            "there is no CPU fallback for the prover hot path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class KzgError(Exception):
    """A failed C-ABI call; `str(e)` is the library's message (the reference's own strings for protocol errors)."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


def as_ptr(b):
    """void* view of a bytes / bytearray / memoryview / numpy array / integer address (no copy where possible)."""
    if b is None:
        return None
    if isinstance(b, int):
        return C.c_void_p(b)
    if isinstance(b, bytes):
        return C.cast(C.c_char_p(b), C.c_void_p)
    if isinstance(b, bytearray):
        return C.cast((C.c_char * len(b)).from_buffer(b), C.c_void_p)
    if hasattr(b, "ctypes"):  # numpy
        return C.c_void_p(b.ctypes.data)
    if hasattr(b, "data_ptr"):  # torch tensor
        return C.c_void_p(b.data_ptr())
    mv = memoryview(b)
    if mv.readonly:
        return C.cast(C.c_char_p(mv.tobytes()), C.c_void_p)
    return C.cast((C.c_char * mv.nbytes).from_buffer(mv), C.c_void_p)
