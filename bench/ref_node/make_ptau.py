"""Synthetic .ptau files for the Node-side runs (CPU only, no GPU needed): sections 1, 2, 3 from the known tau of seed
1001 -- byte for byte what the GPU tests' device generator writes (tests/test_gpu_large_parity.py pins the two).

    python bench/ref_node/make_ptau.py --out-dir tmp [--powers 6,9,11,16]

G1 points come from the threaded C oracle (oracle/c, fixed-base products), the two G2 points from the Python oracle.
Power p holds 2^(p+1) G1 points (one more than snarkjs' 2^(p+1) - 1, so the reference's 2n-point read at n = 2^p stays
inside section 2)."""
import argparse
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.c import binding as oc  # noqa: E402
from oracle.py import bn254 as bn, inputs  # noqa: E402

TAU_SEED = 1001


def write(path, power, tau):
    n_g1 = 1 << (power + 1)
    hdr = struct.pack("<I", 32) + bn.Q.to_bytes(32, "little") + struct.pack("<II", power, power)
    pts = oc.srs_generate(tau, n_g1)
    g2 = bn.g2_to_bytes(bn.G2_GEN) + bn.g2_to_bytes(bn.g2_mul(bn.G2_GEN, tau))
    with open(path, "wb") as f:
        f.write(b"ptau" + struct.pack("<II", 1, 3))
        for sid, payload in ((1, hdr), (2, pts), (3, g2)):
            f.write(struct.pack("<IQ", sid, len(payload)))
            f.write(payload)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out-dir", default="tmp")
    ap.add_argument("--powers", default="6,9,11,16")
    a = ap.parse_args()
    os.makedirs(a.out_dir, exist_ok=True)
    tau = inputs.tau_from_seed(TAU_SEED)
    for p in [int(x) for x in a.powers.split(",")]:
        path = os.path.join(a.out_dir, "synthetic_%02d.ptau" % p)
        write(path, p, tau)
        print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
