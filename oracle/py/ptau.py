"""`.ptau` reader / synthetic writer and the Keccak transcript -- ORACLE, test infrastructure only.

Restates: reference src/ptau_utils.js:3-24 (header), the binfileutils container it is read through
(@iden3/binfileutils@0.0.11, un-vendored: magic, version, nSections, then id u32 | size u64 | payload),
the prover's SRS read (grandsum/mset_eq_kzg_prover.js:83-85) and the verifier's [tau]_2 read
(grandsum/mset_eq_kzg_verifier.js:18-19), and src/Keccak256Transcript.js:7-52.

The Hermez ceremony file cannot be downloaded offline, so `write_ptau` produces a file with the same
layout from a known tau (sections 1, 2, 3 only -- the only ones the reference touches).
"""
import struct

from . import bn254 as bn
from .keccak import keccak256


def write_ptau(path, power, tau, n_g1=None):
    """Write a synthetic ptau: section 1 header, section 2 tauG1 (2^(power+1) points: one more than
    snarkjs' 2^(power+1)-1 so the reference's 2n-point read at n = 2^power stays inside the
    section), section 3 tauG2 ([1]_2, [tau]_2)."""
    if n_g1 is None:
        n_g1 = 1 << (power + 1)
    hdr = struct.pack("<I", 32) + bn.Q.to_bytes(32, "little") + struct.pack("<II", power, power)
    pts = bytearray()
    # [tau^i]_1 = (tau^i mod r) * G1 through a fixed-base window table; fine up to power ~14 in
    # Python.  Larger files are made by the product's device SRS generator (tests compare a prefix).
    t = 1
    for _ in range(n_g1):
        pts += bn.g1_to_bytes(bn.g1_mul_gen(t))
        t = t * tau % bn.R
    g2 = bn.g2_to_bytes(bn.G2_GEN) + bn.g2_to_bytes(bn.g2_mul(bn.G2_GEN, tau))
    with open(path, "wb") as f:
        f.write(b"ptau" + struct.pack("<II", 1, 3))
        for sid, payload in ((1, hdr), (2, bytes(pts)), (3, g2)):
            f.write(struct.pack("<IQ", sid, len(payload)))
            f.write(payload)


def read_sections(path):
    """binfileutils.readBinFile: returns {id: [(offset, size), ...]}; checks magic and version <= 1."""
    with open(path, "rb") as f:
        data = f.read(12)
        if data[:4] != b"ptau":
            raise ValueError(path + ": Invalid File format")
        version, nsec = struct.unpack("<II", data[4:12])
        if version > 1:
            raise ValueError("Version not supported")
        sections = {}
        pos = 12
        for _ in range(nsec):
            f.seek(pos)
            sid, size = struct.unpack("<IQ", f.read(12))
            sections.setdefault(sid, []).append((pos + 12, size))
            pos += 12 + size
    return sections


def read_ptau_header(path, sections):
    """src/ptau_utils.js:3-24."""
    if 1 not in sections:
        raise ValueError(path + ": File has no  header")
    if len(sections[1]) > 1:
        raise ValueError(path + ": File has more than one header")
    off, size = sections[1][0]
    with open(path, "rb") as f:
        f.seek(off)
        n8 = struct.unpack("<I", f.read(4))[0]
        q = int.from_bytes(f.read(n8), "little")
        if q != bn.Q:
            raise ValueError("Curve not supported")
        if n8 != 32:
            raise ValueError(path + ": Invalid size")
        power, ceremony_power = struct.unpack("<II", f.read(8))
        if 4 + n8 + 8 != size:
            raise ValueError("Invalid PTau header size")
    return power, ceremony_power


def read_tau_g1(path, sections, n_points):
    """prover.js:83-85: n_points * 64 B straight from the start of section 2 (Montgomery-LE affine)."""
    off, _ = sections[2][0]
    with open(path, "rb") as f:
        f.seek(off)
        return f.read(n_points * 64)


def read_tau_g2(path, sections):
    """verifier.js:18-19: the SECOND G2 point of section 3."""
    off, _ = sections[3][0]
    with open(path, "rb") as f:
        f.seek(off + 128)
        return f.read(128)


class Keccak256Transcript:
    """src/Keccak256Transcript.js:7-52 -- every getChallenge() hashes ALL data appended so far."""

    def __init__(self):
        self.data = []

    def reset(self):
        self.data = []

    def add_pol_commitment(self, c64):
        assert len(c64) == 64
        self.data.append((0, bytes(c64)))

    def add_field_element(self, e32):
        assert len(e32) == 32
        self.data.append((1, bytes(e32)))

    def get_challenge(self):
        """returns the challenge as 32 B Montgomery-LE (Fr.e(bigint) converts to Montgomery)."""
        if not self.data:
            raise ValueError("Keccak256Transcript: No data to generate a transcript")
        buf = b""
        for typ, d in self.data:
            buf += bn.g1_to_rpr_uncompressed(d) if typ == 0 else bn.fr_to_rpr_be(d)
        value = int.from_bytes(keccak256(buf), "big")
        return bn.fr_to_mont_bytes(value % bn.R)
