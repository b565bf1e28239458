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


def readPTauHost(pTauFilename):
    """Host-only reader for the verifiers (verifier.js:12-20): binfileutils section scan, readPTauHeader checks with
    the reference's messages, and [tau]_2 = the second G2 point of section 3.  -> {power, ceremonyPower, X2}"""
    import struct
    with open(pTauFilename, "rb") as f:
        head = f.read(12)
        if len(head) < 12 or head[:4] != b"ptau":
            raise ValueError(pTauFilename + ": Invalid File format")
        version, nsec = struct.unpack("<II", head[4:12])
        if version > 1:
            raise ValueError("Version not supported")
        sections = {}
        pos = 12
        for _ in range(nsec):
            f.seek(pos)
            sh = f.read(12)
            if len(sh) < 12:
                raise ValueError(pTauFilename + ": truncated section table")
            sid, size = struct.unpack("<IQ", sh)
            sections.setdefault(sid, []).append((pos + 12, size))
            pos += 12 + size
        if 1 not in sections:
            raise ValueError(pTauFilename + ": File has no  header")
        if len(sections[1]) > 1:
            raise ValueError(pTauFilename + ": File has more than one header")
        off, size = sections[1][0]
        f.seek(off)
        n8 = struct.unpack("<I", f.read(4))[0]
        q = int.from_bytes(f.read(n8), "little")
        from .curve import Q
        if q != Q:
            raise ValueError("Curve not supported")
        if n8 != 32:
            raise ValueError(pTauFilename + ": Invalid size")
        power, ceremony = struct.unpack("<II", f.read(8))
        if 4 + n8 + 8 != size:
            raise ValueError("Invalid PTau header size")
        if 3 not in sections or sections[3][0][1] < 256:
            raise ValueError(pTauFilename + ": no tauG2 section")
        f.seek(sections[3][0][0] + 128)
        x2 = f.read(128)
    return {"power": power, "ceremonyPower": ceremony, "X2": x2}
