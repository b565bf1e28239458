"""Generate the committed golden proofs from the CPU oracle (run here, CPU only):
    python tests/golden/make_golden.py
Inputs are the seeded synthetic columns of tests/test_gpu_prover.py::make_case; the SRS is the synthetic
tau of seed 1001.  Each fixture stores the canonical proof bytes (raw values in key order)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.py import bn254 as bn, inputs, protocol as pr  # noqa: E402

TAU_SEED = 1001
CASES = [
    # name, kind, seed, nbits, k, selected, rotate, ptau_power
    ("gs_c1_n8", "gs", 1, 8, 1, False, True, 11),
    ("gp_c1_n8", "gp", 1, 8, 1, False, True, 11),
    ("gs_c2_n11", "gs", 2, 11, 1, False, True, 11),
    ("gp_c2_n11", "gp", 2, 11, 1, False, True, 11),
    ("gs_vec_sel_n6_k3", "gs", 21, 6, 3, True, False, 6),
    ("gp_vec_sel_n6_k3", "gp", 21, 6, 3, True, False, 6),
    ("gs_perm_n9", "gs", 31, 9, 1, False, False, 9),
    ("gp_perm_n9", "gp", 31, 9, 1, False, False, 9),
]


def columns(seed, nbits, k, selected, rotate):
    n = 1 << nbits
    cols_f = [inputs.random_column(seed * 100 + i, n) for i in range(k)]
    if selected or rotate:
        cols_t = [inputs.rotate_right(c) for c in cols_f]
    else:
        perm = inputs.permutation(seed, n)
        cols_t = [[c[perm[i]] for i in range(n)] for c in cols_f]
    sel_f = sel_t = None
    if selected:
        one, zero = bn.fr_to_mont_bytes(1), bytes(32)
        sel_f = one * (n - 1) + zero
        sel_t = zero + one * (n - 1)
    return cols_f, cols_t, sel_f, sel_t


def main():
    tau = inputs.tau_from_seed(TAU_SEED)
    for name, kind, seed, nbits, k, selected, rotate, power in CASES:
        cf, ct, sf, st = columns(seed, nbits, k, selected, rotate)
        prover = pr.grandsum_prover if kind == "gs" else pr.grandproduct_prover
        verifier = pr.grandsum_verifier if kind == "gs" else pr.grandproduct_verifier
        trace = {}
        proof = prover(pr.TrapdoorSrs(tau, power), [bn.fr_vec_to_std_bytes(c) for c in cf],
                       [bn.fr_vec_to_std_bytes(c) for c in ct], sf, st, trace=trace)
        assert verifier(proof, nbits, tau=tau)
        out = {
            "kind": kind, "seed": seed, "nbits": nbits, "k": k, "selected": selected, "rotate": rotate,
            "ptau_power": power, "tau_seed": TAU_SEED,
            "commitment_keys": list(proof["commitments"]), "evaluation_keys": list(proof["evaluations"]),
            "challenges": {k_: hex(v) for k_, v in trace["challenges"].items()},
            "proof_bytes": pr.proof_bytes(proof).hex(),
        }
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(out, f, indent=1)
        print(name, "ok")


if __name__ == "__main__":
    main()
