"""In-tree build of the C-ABI library (nvcc, sm_100a only).

Produces lambda_snark_r_b200/lib/liblambda_snark_core.so (what ctypes / a C
harness loads) and liblambda_snark_core.a (what lambda-snark-sys would link;
same name as the reference's static library, cpp-core/CMakeLists.txt:107).
Objects are rebuilt only when a source or header is newer.

`python -m lambda_snark_r_b200._build --profiling` builds a SECOND library,
lib/liblambda_snark_core_prof.so, with -DLSR_PROFILING: the only build in which
the phase-skip / kernel-selection switches of tools/ exist (LSR_FUSED_SKIP,
LSR_NTT_COLUMN2, LSR_FS_KERNEL).  The shipped library reads no environment.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OBJ = PKG / "_build"
LIB = PKG / "lib"
SO = LIB / "liblambda_snark_core.so"
AR = LIB / "liblambda_snark_core.a"

SOURCES = ["lsr_host.cpp", "lsr_r1cs.cpp", "lsr_abi.cpp", "lsr_ntt.cu", "lsr_commit.cu", "lsr_commit_fused.cu",
           "lsr_microbench.cu", "lsr_quotient.cu", "lsr_fiat_shamir.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _flags() -> list[str]:
    return ["-std=c++17", "-O3", "-lineinfo", *ARCH, "-Xcompiler", "-fPIC,-Wall,-Wno-unknown-pragmas",
            "-ccbin", shutil.which("g++") or "g++",
            "-I", str(CSRC), "-I", str(ROOT / "include")]


def _newest_header() -> float:
    hs = list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list((ROOT / "include").rglob("*.h"))
    return max(h.stat().st_mtime for h in hs)


def _compile(src: str, force: bool, verbose: bool, objdir: Path = OBJ, extra: tuple = ()) -> Path:
    s = CSRC / src
    o = objdir / (s.stem + ".o")
    stamp = max(s.stat().st_mtime, _newest_header(), Path(__file__).stat().st_mtime)
    if not force and o.exists() and o.stat().st_mtime >= stamp:
        return o
    cmd = [_nvcc(), *_flags(), *extra, "-x", "cu" if s.suffix == ".cu" else "c++", "-c", str(s), "-o", str(o)]
    if s.suffix != ".cu":
        cmd = [_nvcc(), *_flags(), *extra, "-c", str(s), "-o", str(o)]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose and r.stderr.strip():
        print(r.stderr)
    return o


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    LIB.mkdir(exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(6, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose), SOURCES))
    newest = max(o.stat().st_mtime for o in objs)
    if force or not SO.exists() or SO.stat().st_mtime < newest:
        cmd = [_nvcc(), "-shared", *ARCH, "-ccbin", shutil.which("g++") or "g++", "-cudart", "static",
               "-o", str(SO), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
        if AR.exists():
            AR.unlink()
        subprocess.run(["ar", "rcs", str(AR), *map(str, objs)], check=True)
    return SO


def build_profiling(force: bool = False, verbose: bool = False) -> Path:
    """tools/ only: the library with the profiling switches compiled in (never loaded by the package itself)."""
    objdir = PKG / "_build_prof"
    objdir.mkdir(exist_ok=True)
    LIB.mkdir(exist_ok=True)
    so = LIB / "liblambda_snark_core_prof.so"
    with ThreadPoolExecutor(max_workers=min(6, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose, objdir, ("-DLSR_PROFILING",)), SOURCES))
    if force or not so.exists() or so.stat().st_mtime < max(o.stat().st_mtime for o in objs):
        cmd = [_nvcc(), "-shared", *ARCH, "-ccbin", shutil.which("g++") or "g++", "-cudart", "static",
               "-o", str(so), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return so


def build_variant(name: str, defines: tuple, force: bool = False, verbose: bool = False) -> Path:
    """tools/ only: lib/liblambda_snark_core_<name>.so compiled with extra -D switches (kernel-variant experiments;
    never loaded by the package itself)."""
    objdir = PKG / f"_build_{name}"
    objdir.mkdir(exist_ok=True)
    LIB.mkdir(exist_ok=True)
    so = LIB / f"liblambda_snark_core_{name}.so"
    with ThreadPoolExecutor(max_workers=min(6, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose, objdir, tuple(defines)), SOURCES))
    if force or not so.exists() or so.stat().st_mtime < max(o.stat().st_mtime for o in objs):
        cmd = [_nvcc(), "-shared", *ARCH, "-ccbin", shutil.which("g++") or "g++", "-cudart", "static",
               "-o", str(so), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return so


if __name__ == "__main__":
    if "--variant" in sys.argv:      # python -m lambda_snark_r_b200._build --variant NAME -DX=1 -DY=2
        i = sys.argv.index("--variant")
        print("built", build_variant(sys.argv[i + 1], tuple(a for a in sys.argv[i + 2:] if a.startswith("-D")), verbose=True))
        sys.exit(0)
    if "--profiling" in sys.argv:
        print("built", build_profiling(force="--force" in sys.argv, verbose=True))
    else:
        print("built", build(force="--force" in sys.argv, verbose=True))
