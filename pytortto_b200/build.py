"""Builds libtortto_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m pytortto_b200.build [--force] [--tuning]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  `--tuning` additionally builds
libtortto_b200_tuning.so with -DTTB_TUNING: the only build whose kernels read TTB_* experiment variables from the
environment (selected at run time with TORTTO_B200_LIB=tuning; never loaded by default).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libtortto_b200.so")
OBJ = os.path.join(HERE, "_build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--extended-lambda", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
    "-Xptxas", "-v", "-I", INCLUDE,
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path, extra=()):
    h = hashlib.sha1()
    h.update(" ".join(list(NVCC_FLAGS) + list(extra)).encode())
    for p in [path] + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + \
            [os.path.join(INCLUDE, "tortto_b200.h")]:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False, tuning=False):
    obj_dir = OBJ + ("_tuning" if tuning else "")
    lib_path = LIB.replace(".so", "_tuning.so") if tuning else LIB
    extra = ["-DTTB_TUNING"] if tuning else []
    os.makedirs(obj_dir, exist_ok=True)
    objs, rebuilt = [], False
    procs = []
    for src in sources():
        base = os.path.splitext(os.path.basename(src))[0]
        obj = os.path.join(obj_dir, base + ".o")
        stamp = obj + ".sha1"
        dig = _digest(src, extra)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-c", src, "-o", obj]
        procs.append((src, stamp, dig, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, stamp, dig, p in procs:
        out, _ = p.communicate()
        log = os.path.join(obj_dir, os.path.basename(src) + ".log")
        with open(log, "w") as f:
            f.write(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src}")
        if verbose:
            sys.stdout.write(out)
        with open(stamp, "w") as f:
            f.write(dig)
        rebuilt = True
    if rebuilt or force or not os.path.exists(lib_path):
        cmd = [_nvcc(), "-shared", "-o", lib_path] + objs  # static cudart; driver entry points are resolved at run time
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--tuning" in sys.argv:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, tuning=True))
