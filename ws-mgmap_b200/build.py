"""In-tree build of the native libraries (no torch involved, plain nvcc / g++).

    python ws-mgmap_b200/build.py            # libwsmg.so (sm_100a) + test-only emulation
The .so files land in ws-mgmap_b200/lib/ (git-ignored, shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")
ROOT = os.path.dirname(HERE)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                 # FMAs only where the arithmetic contract writes fmaf()
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared", "-ldl",
]


def _digest(sources, extra="") -> str:
    import hashlib
    h = hashlib.sha256(extra.encode())
    for s in sources:
        with open(s, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _newer(target: str, sources, extra="") -> bool:
    """True when `target` was built from exactly these sources (content hash in target + '.stamp'; file times do
    not survive the copy to the GPU box)."""
    stamp = target + ".stamp"
    if not (os.path.exists(target) and os.path.exists(stamp)):
        return False
    with open(stamp) as f:
        return f.read().strip() == _digest(sources, extra)


def _stamp(target: str, sources, extra=""):
    with open(target + ".stamp", "w") as f:
        f.write(_digest(sources, extra))


def _sources():
    out = [os.path.join(ROOT, "include", "wsmg.h")]
    for f in sorted(os.listdir(CSRC)):
        out.append(os.path.join(CSRC, f))
    return out


def build_cuda(verbose: bool = False, force: bool = False, phase_skip: bool = False, variant: str = "", defines=(), flags=()) -> str:
    """phase_skip=True builds the profiling variant lib/libwsmg_phaseskip.so (-DWSMG_PHASE_SKIP, see
    csrc/wsmg_body.h: WSMG_SKIP); load it with WSMG_LIB_PATH, never in production.  `variant` + `defines` build an
    experiment library lib/libwsmg_<variant>.so with extra -D switches (scripts/variants.py)."""
    os.makedirs(LIB, exist_ok=True)
    target = os.path.join(LIB, f"libwsmg_{variant}.so" if variant else ("libwsmg_phaseskip.so" if phase_skip else "libwsmg.so"))
    if not force and _newer(target, _sources()):
        return target
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-o", target, os.path.join(CSRC, "wsmg.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    for d in defines:
        cmd.insert(1, "-D" + d)
    for f in flags:
        cmd.insert(1, f)
    if phase_skip:
        cmd.insert(1, "-DWSMG_PHASE_SKIP")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    _stamp(target, _sources())
    return target


def build_emulation(force: bool = False) -> str:
    """Host emulation of the kernels -- test infrastructure, see csrc/wsmg_emul.cpp."""
    os.makedirs(LIB, exist_ok=True)
    target = os.path.join(LIB, "libwsmg_emul.so")
    if not force and _newer(target, _sources()):
        return target
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-mfma", "-fno-fast-math",
           "-I", os.path.join(ROOT, "include"), "-o", target, os.path.join(CSRC, "wsmg_emul.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    _stamp(target, _sources())
    return target


if __name__ == "__main__":
    v = "-v" in sys.argv
    print(build_cuda(verbose=v, force="-f" in sys.argv))
    print(build_emulation(force="-f" in sys.argv))
    if "--phase-skip" in sys.argv:
        print(build_cuda(force=True, phase_skip=True))
