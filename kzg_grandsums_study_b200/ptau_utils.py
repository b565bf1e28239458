"""readPTauHeader -- mirror of reference src/ptau_utils.js:3-24 (plus the readBinFile section scan of
@iden3/binfileutils it sits on, prover.js:15).  The header checks run in libkzgb200.so; section 2 is
uploaded to the device by curve.load_srs()."""
import ctypes as C

from .curve import getCurveFromName


def readPTauHeader(pTauFilename, device=0):
    """-> {curve, power, ceremonyPower}.  Raises with the reference's messages on a malformed file."""
    curve = getCurveFromName("bn128", device)
    power = C.c_uint32()
    ceremony = C.c_uint32()
    curve.check(curve.lib.kzg_ptau_read_header(curve.ctx, pTauFilename.encode(), C.byref(power), C.byref(ceremony)))
    return {"curve": curve, "power": power.value, "ceremonyPower": ceremony.value}


def readTauG2(pTauFilename, curve):
    """[tau]_2: the second G2 point of section 3 (verifier.js:18-19), 128 bytes"""
    out = bytearray(128)
    from ._lib import as_ptr
    curve.check(curve.lib.kzg_ptau_read_tau_g2(curve.ctx, pTauFilename.encode(), as_ptr(out)))
    return bytes(out)
