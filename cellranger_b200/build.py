"""Builds libcrgpu.so (the CUDA kernels + the C ABI of include/crgpu.h) for sm_100a, in-tree."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libcrgpu.so")
SOURCES = ["crgpu.cu", "shard.cu", "pass_kernels.cu", "sort.cu", "dedup_kernels.cu", "finish_kernels.cu", "compat.cu", "synth.cu", "fastq.cu", "mex_writer.cpp"]
HEADERS = ["common.cuh", "kernels.h", "ctx.h", "nccl_dl.h", os.path.join(ROOT, "include", "crgpu.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libcrgpu.so cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile libcrgpu.so if it is missing or older than its sources. Several processes may get here at once
    (every torchrun rank imports the package): the build runs under an exclusive file lock, into a temporary
    file that replaces the library atomically, so nobody ever loads a half-written file."""
    if not force and not is_stale():
        return SO
    import fcntl

    with open(SO + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():  # another process built it while we waited
                return SO
            tmp = f"{SO}.tmp.{os.getpid()}"
            # one nvcc per translation unit, in parallel; then one link
            objs, procs = [], []
            for s in SOURCES:
                obj = os.path.join(CSRC, f".{s}.{os.getpid()}.o")
                objs.append(obj)
                cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else []) + \
                      ["-c", os.path.join(CSRC, s), "-o", obj]
                procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
            logs = [p.communicate()[0] for p in procs]
            try:
                if any(p.returncode != 0 for p in procs):
                    raise RuntimeError("nvcc failed:\n" + "\n".join(logs))
                res = subprocess.run([_nvcc(), "-shared", "-o", tmp] + objs + ["-lz", "-ldl"], capture_output=True, text=True)
                if res.returncode != 0:
                    raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
                os.replace(tmp, SO)
            finally:
                for f in objs + [tmp]:
                    if os.path.exists(f):
                        os.unlink(f)
            if verbose:
                print("\n".join(logs))
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return SO


if __name__ == "__main__":
    import sys

    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
