"""Builds libsoftbody_b200.so in-tree with nvcc for sm_100a (no JIT cache: the
built file travels to the GPU box with the repo snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsoftbody_b200.so")
SOURCES = ["solver.cu", "plan.cpp", "ingest.cpp"]
DEPS = SOURCES + ["kernels.cuh", "plan.h", os.path.join("..", "..", "include", "softbody_b200.h")]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -ffp-contract=off: host code (rest values, step constants) must round exactly like
# the arithmetic contract says; device code uses explicit rounding intrinsics.
HOST_FLAGS = "-fPIC,-ffp-contract=off,-march=x86-64-v3,-O2,-Wall,-pthread"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-ccbin", "/usr/bin/g++", "-Xcompiler", HOST_FLAGS, "-shared", "-cudart", "static",
]


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles the library if it is stale.  Safe with several processes at once (torchrun: every rank imports the
    package): one of them builds under a file lock, into a temporary file that is renamed into place; the others wait
    for the lock and find the library fresh."""
    if not force and not stale():
        return LIB
    import fcntl
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not stale():
                return LIB  # another process built it while this one waited
            tmp = f"{LIB}.tmp.{os.getpid()}"
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
                ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lpthread"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed building libsoftbody_b200.so")
            os.replace(tmp, LIB)
            if verbose:
                sys.stderr.write(r.stdout + r.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
