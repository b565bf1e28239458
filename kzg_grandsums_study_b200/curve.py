"""The `curve` object of the reference's host code, backed by libkzgb200.so.

Mirrors what ffjavascript's `getCurveFromName("bn128")` / `getCurveFromQ(q)` hands to the reference
(ptau_utils.js:13; test/mset_eq_kzg_grandsum.test.js:14-21): `curve.Fr` scalar-field helpers on 32-byte
Montgomery-LE elements, `curve.G1` sizes, `curve.terminate()`.  Bulk work goes to the device through
the C ABI; only O(1) scalar conversions are done with Python integers on the host.
"""
import ctypes as C
import os

from . import _lib
from ._lib import KzgError, as_ptr

Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_MONT = 1 << 256
_MONT_INV_R = pow(_MONT, -1, R)


class DeviceBuffer:
    """A device-resident vector of Fr elements (the `.coef` / `.eval` BigBuffer of the reference)."""

    def __init__(self, curve, handle):
        self.curve = curve
        self.handle = handle

    @property
    def byteLength(self):
        return self.length() * 32

    def length(self):
        return int(self.curve.lib.kzg_buf_len(self.handle)) if self.handle else 0

    def tobytes(self):
        n = self.length()
        out = bytearray(32 * n)
        if n:
            self.curve.check(self.curve.lib.kzg_buf_download(self.curve.ctx, self.handle, 0, as_ptr(out), n))
        return bytes(out)

    def slice(self, start=0, end=None):
        """byte-offset slice like Uint8Array.slice; returns host bytes"""
        end = self.byteLength if end is None else end
        if start % 32 or end % 32:
            return self.tobytes()[start:end]
        n = (end - start) // 32
        out = bytearray(32 * n)
        if n:
            self.curve.check(self.curve.lib.kzg_buf_download(self.curve.ctx, self.handle, start // 32, as_ptr(out), n))
        return bytes(out)

    def clone(self):
        out = self.curve.alloc(self.length())
        self.curve.check(self.curve.lib.kzg_buf_copy(self.curve.ctx, out.handle, 0, self.handle, 0, self.length()))
        return out

    def free(self):
        if self.handle and self.curve.ctx:
            self.curve.lib.kzg_buf_free(self.curve.ctx, self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _Fr:
    """curve.Fr: 32-byte little-endian Montgomery elements."""
    n8 = 32
    p = R

    def __init__(self, curve):
        self._curve = curve
        self.zero = bytes(32)
        self.one = self.e(1)
        self.negone = self.e(R - 1)
        # Fr.w[k] = 5^((r-1)/2^k), Montgomery
        w = [0] * 29
        w[28] = pow(5, (R - 1) >> 28, R)
        for i in range(27, -1, -1):
            w[i] = w[i + 1] * w[i + 1] % R
        self.w = [self.e(x) for x in w]

    def e(self, x):
        if isinstance(x, (bytes, bytearray, memoryview)):
            return bytes(x)
        return ((int(x) % R) * _MONT % R).to_bytes(32, "little")

    def toObject(self, b):
        return int.from_bytes(bytes(b), "little") * _MONT_INV_R % R

    def toString(self, b, radix=10):
        v = self.toObject(b)
        return str(v) if radix == 10 else format(v, "x")

    def _bin(self, a, b, op):
        return self.e(op(self.toObject(a), self.toObject(b)))

    def add(self, a, b):
        return self._bin(a, b, lambda x, y: x + y)

    def sub(self, a, b):
        return self._bin(a, b, lambda x, y: x - y)

    def mul(self, a, b):
        return self._bin(a, b, lambda x, y: x * y)

    def neg(self, a):
        return self.e(-self.toObject(a))

    def square(self, a):
        return self.mul(a, a)

    def inv(self, a):
        return self.e(pow(self.toObject(a), -1, R))

    def div(self, a, b):
        return self.mul(a, self.inv(b))

    def eq(self, a, b):
        return bytes(a) == bytes(b)

    def isZero(self, a):
        return bytes(a) == self.zero

    def exp(self, a, k):
        return self.e(pow(self.toObject(a), int(k), R))

    def random(self):
        """raw LE bytes of a uniform value < r -- NOT Montgomery-converted, exactly like ffjavascript's Fr.random()"""
        return (int.from_bytes(os.urandom(48), "little") % R).to_bytes(32, "little")

    def toRprBE(self, b):
        out = bytearray(32)
        self._curve.lib.kzg_fr_to_rpr_be(as_ptr(bytes(b)), as_ptr(out))
        return bytes(out)

    # bulk calls: device
    def batchToMontgomery(self, buf):
        return self._curve._convert(buf, True)

    def batchFromMontgomery(self, buf):
        return self._curve._convert(buf, False)

    def fft(self, buf):
        return self._curve._ntt(buf, False)

    def ifft(self, buf):
        return self._curve._ntt(buf, True)

    def batchInverse(self, buf):
        d = self._curve.to_device(buf)
        out = self._curve.alloc(d.length(), zero=False)
        self._curve.check(self._curve.lib.kzg_fr_batch_inverse(self._curve.ctx, d.handle, out.handle))
        return out


class _F1:
    n8 = 32
    n64 = 4


class _G1:
    F = _F1()

    def __init__(self, curve):
        self._curve = curve
        self.zero = bytes(64)

    def toRprUncompressed(self, p64):
        out = bytearray(64)
        self._curve.lib.kzg_g1_to_rpr_uncompressed(as_ptr(bytes(p64)), as_ptr(out))
        return bytes(out)

    def toObject(self, p64):
        """affine point as a pair of integers (None for infinity)"""
        b = bytes(p64)
        if b == self.zero:
            return None
        inv = pow(_MONT, -1, Q)
        return (int.from_bytes(b[:32], "little") * inv % Q, int.from_bytes(b[32:], "little") * inv % Q)

    def multiExpAffine(self, bases, scalars):
        """G1.multiExpAffine(bases 64 B affine Montgomery-LE, scalars 32 B standard-form LE) -> 96 B Jacobian"""
        n = len(scalars) // 32
        aff = bytearray(64)
        jac = bytearray(96)
        self._curve.check(self._curve.lib.kzg_g1_msm_affine(self._curve.ctx, as_ptr(bytes(bases)), as_ptr(bytes(scalars)), n, 0,
                                                            as_ptr(aff), as_ptr(jac)))
        return bytes(jac)

    def toAffine(self, jac):
        return bytes(jac[:64])


class HostCurve:
    """the host-only part of the curve object (Fr / G1 byte helpers, Keccak): what the verifiers and the transcript
    need.  No device, no context."""
    name = "bn128"
    q = Q
    r = R

    def __init__(self):
        self.lib = _lib.load()
        self.ctx = None
        self.Fr = _Fr(self)
        self.G1 = _G1(self)
        self.F1 = _F1()

    def terminate(self):
        pass


_HOST_CURVE = None


def getHostCurve():
    global _HOST_CURVE
    if _HOST_CURVE is None:
        _HOST_CURVE = HostCurve()
    return _HOST_CURVE


class Curve(HostCurve):
    def __init__(self, device=0, stream=None):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.kzg_ctx_create(device, stream, C.byref(h))
        if rc != 0:
            raise KzgError(rc, "kzg_ctx_create failed (code %d): a CUDA device is required, there is no CPU fallback" % rc)
        self.ctx = h
        self.device = device
        self.Fr = _Fr(self)
        self.G1 = _G1(self)
        self.F1 = _F1()
        self._srs_cache = {}

    # ---- plumbing -----------------------------------------------------------------------------------
    def check(self, rc):
        if rc != 0:
            msg = self.lib.kzg_last_error(self.ctx)
            raise KzgError(rc, msg.decode() if msg else "kzg error %d" % rc)

    def alloc(self, n, zero=True):
        """a zero-filled device vector of n elements (kzg_buf_alloc always clears, like `new Uint8Array`)"""
        h = C.c_void_p()
        self.check(self.lib.kzg_buf_alloc(self.ctx, n, C.byref(h)))
        return DeviceBuffer(self, h)

    def to_device(self, data):
        """host bytes (or an existing DeviceBuffer) -> DeviceBuffer"""
        if isinstance(data, DeviceBuffer):
            return data
        data = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
        if len(data) % 32:
            raise ValueError("buffer length is not a multiple of 32")
        n = len(data) // 32
        buf = self.alloc(n)
        if n:
            self.check(self.lib.kzg_buf_upload(self.ctx, buf.handle, 0, as_ptr(data), n))
        return buf

    def wrap(self, handle):
        return DeviceBuffer(self, handle)

    def sync(self):
        self.check(self.lib.kzg_ctx_sync(self.ctx))

    def launch_count(self):
        return int(self.lib.kzg_ctx_launch_count(self.ctx))

    def set_option(self, name, value):
        """MSM tuning knob (kzg_ctx_set_option): A/B timing and the forced paths of the tests; value < 0 = default"""
        self.check(self.lib.kzg_ctx_set_option(self.ctx, name.encode(), int(value)))

    def _convert(self, buf, to_mont):
        d = self.to_device(buf)
        out = self.alloc(d.length(), zero=False)
        fn = self.lib.kzg_fr_to_mont if to_mont else self.lib.kzg_fr_from_mont
        self.check(fn(self.ctx, d.handle, out.handle))
        return out

    def _ntt(self, buf, inverse):
        d = self.to_device(buf)
        n = d.length()
        if n == 0 or n & (n - 1):
            raise KzgError(_lib.KZG_ERR_PROTOCOL, "fft must be multiple of 2")
        out = self.alloc(n, zero=False)
        self.check(self.lib.kzg_fr_ntt(self.ctx, d.handle, out.handle, 1 if inverse else 0))
        return out

    # ---- SRS ----------------------------------------------------------------------------------------
    def load_srs(self, ptau_path, n_points):
        """device-resident [tau^i]_1 from a .ptau file (prover.js:15-16,83-85).  Cached per (path, size AND the file's
        identity: mtime, byte size) -- a .ptau rewritten in place with another tau is loaded afresh, the stale device
        copy (and its window table) is released."""
        path = os.path.abspath(ptau_path)
        st = os.stat(path)
        key = (path, int(n_points))
        ident = (st.st_mtime_ns, st.st_size)
        hit = self._srs_cache.get(key)
        if hit is not None and hit[2] != ident:
            self.lib.kzg_srs_free(self.ctx, hit[0])
            del self._srs_cache[key]
            hit = None
        if hit is None:
            h = C.c_void_p()
            power = C.c_uint32()
            self.check(self.lib.kzg_srs_load_ptau(self.ctx, ptau_path.encode(), n_points, C.byref(h), C.byref(power)))
            if os.environ.get("KZGB200_NO_SRS_TABLE") != "1":
                c = int(os.environ.get("KZGB200_TABLE_WINDOW", "0"))        # 0 = the library's cost model
                self.check(self.lib.kzg_srs_precompute(self.ctx, h, c))   # one-off window table (msm.cu)
            hit = self._srs_cache[key] = (h, power.value, ident)
        return hit[0], hit[1]

    def load_lagrange_srs(self, ptau_path, nBits):
        """[L_i(tau)]_1, i < 2^nBits, derived on the device from the monomial points of a .ptau (kzg_srs_lagrange; a
        one-off per (file, size), cached like load_srs).  `Evaluations.commit(lagrange)` / kzg_commit over it commits a
        polynomial given by its evaluations on H -- same point as committing its iNTT over the monomial SRS."""
        n = 1 << nBits
        mono, _ = self.load_srs(ptau_path, n)
        path = os.path.abspath(ptau_path)
        st = os.stat(path)
        key = (path, "lagrange", int(nBits))
        ident = (st.st_mtime_ns, st.st_size)
        hit = self._srs_cache.get(key)
        if hit is not None and hit[2] != ident:
            self.lib.kzg_srs_free(self.ctx, hit[0])
            hit = None
        if hit is None:
            h = C.c_void_p()
            self.check(self.lib.kzg_srs_lagrange(self.ctx, mono, nBits, C.byref(h)))
            self.check(self.lib.kzg_srs_precompute(self.ctx, h, 0))
            hit = self._srs_cache[key] = (h, nBits, ident)
        return hit[0]

    def terminate(self):
        if self.ctx:
            for entry in self._srs_cache.values():
                self.lib.kzg_srs_free(self.ctx, entry[0])
            self._srs_cache = {}
            self.lib.kzg_ctx_destroy(self.ctx)
            self.ctx = None
        if _CURVES.get(self.device) is self:
            del _CURVES[self.device]


_CURVES = {}


def getCurveFromName(name="bn128", device=0):
    """ffjavascript getCurveFromName: the curve object is cached until curve.terminate()"""
    if name.lower() not in ("bn128", "bn254", "altbn128"):
        raise ValueError("Curve not supported: %s" % name)
    if device not in _CURVES:
        _CURVES[device] = Curve(device)
    return _CURVES[device]


def getCurveFromQ(q, device=0):
    if int(q) != Q:
        raise ValueError("Curve not supported: %s" % q)
    return getCurveFromName("bn128", device)
