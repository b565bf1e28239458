"""Keccak256Transcript -- mirror of reference src/Keccak256Transcript.js:7-52.  Stays on the HOST by mandate.

An append-only list of tagged items; getChallenge() serialises everything added so far (commitments as
G1.toRprUncompressed, 64 B big-endian standard form; scalars as Fr.toRprBE), hashes with Keccak-256
(original 0x01 padding, js-sha3 `keccak256`) and maps the big-endian digest into Fr (Montgomery).
The hashing and byte conversions are the host helpers of libkzgb200.so (no device work).
"""
from ._lib import as_ptr

POLYNOMIAL = 0
SCALAR = 1


class Keccak256Transcript:
    def __init__(self, curve):                                  # :8-13
        self.curve = curve
        self.G1 = curve.G1
        self.Fr = curve.Fr
        self.reset()

    def reset(self):                                            # :15-17
        self.data = []

    def addPolCommitment(self, polynomialCommitment):           # :19-21
        self.data.append((POLYNOMIAL, bytes(polynomialCommitment)))

    def addFieldElement(self, scalar):                          # :23-25
        self.data.append((SCALAR, bytes(scalar)))

    def getChallenge(self):                                     # :27-52
        if not self.data:
            raise ValueError("Keccak256Transcript: No data to generate a transcript")
        buf = bytearray()
        for kind, item in self.data:
            if kind == POLYNOMIAL:
                buf += self.G1.toRprUncompressed(item)
            else:
                buf += self.Fr.toRprBE(item)
        lib = self.curve.lib
        digest = bytearray(32)
        lib.kzg_keccak256(as_ptr(bytes(buf)), len(buf), as_ptr(digest))
        out = bytearray(32)
        lib.kzg_fr_from_hash_be(as_ptr(bytes(digest)), as_ptr(out))
        return bytes(out)
