"""CPU check of the field inversion the device uses (csrc/field.cuh is host-compilable): fp_inv (binary GCD, 31 steps per
multi-limb update) == a^(p-2) == the plain binary Euclid loop on 2 x 20 000 seeded values incl. short ones, powers of
two and p - small.  Replaces nothing of the reference by itself: it sits under G1.toAffine and Fr.batchInverse
(reference src/polynomial/polynomial.js:1113, src/grandsum/grandsum.js:41)."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_bingcd_inverse_matches_fermat_on_host(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "host_field_inverse")
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-I", os.path.join(ROOT, "kzg_grandsums_study_b200", "csrc"),
                           "-o", exe, os.path.join(HERE, "host_field_inverse.cpp")])
    out = subprocess.run([exe, "20000"], stdout=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0, out.stdout
    assert "Fq: 20000 cases, 0 bad" in out.stdout and "Fr: 20000 cases, 0 bad" in out.stdout
