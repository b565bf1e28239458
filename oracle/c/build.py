"""Build the oracle's C restatement (test infrastructure): gcc -O3 -fopenmp -shared -> oracle/c/libkzgoracle.so"""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "kzg_oracle.c")
LIB = os.path.join(HERE, "libkzgoracle.so")
FLAGS = ["-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", "-std=gnu11", "-Wall"]


def _cpu_id():
    """-march=native ties the binary to the build host: rebuild when the library travels to a different CPU"""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return hashlib.sha256(line.encode()).hexdigest()[:16]
    except OSError:
        pass
    return "unknown"


def build(verbose=False):
    with open(SRC, "rb") as f:
        digest = hashlib.sha256(f.read() + " ".join(FLAGS).encode() + _cpu_id().encode()).hexdigest()
    stamp = LIB + ".stamp"
    if os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        if verbose:
            print("[oracle build] libkzgoracle.so is up to date")
        return LIB
    flags = list(FLAGS)
    p = subprocess.run(["gcc"] + flags + [SRC, "-o", LIB], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0 and "-march=native" in flags:      # a box whose CPU the compiler does not know
        flags.remove("-march=native")
        p = subprocess.run(["gcc"] + flags + [SRC, "-o", LIB], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        raise RuntimeError("oracle C build failed:\n" + p.stdout)
    if verbose and p.stdout.strip():
        print(p.stdout)
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
