"""Build the native library (CUDA kernels for sm_100a + C++ host runtime) in-tree.

``python -m infomesh_b200.build`` compiles every ``csrc/**/*.cu|*.cpp`` with
``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` into
``infomesh_b200/_native/libinfomesh_b200.so``.  nvcc cross-compiles without a GPU, so this
runs on the CPU-only build box; the resulting ``.so`` travels with the tree to the B200 box.
Objects are rebuilt only when their source (or any header) is newer.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
OUT_DIR = ROOT / "_native"
OBJ_DIR = OUT_DIR / "obj"
LIB_PATH = OUT_DIR / "libinfomesh_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3",
    "-Xptxas", "-v",
    "-DIM_BUILD",
]


def find_nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")
    return cand


def sources() -> list[Path]:
    return sorted([*CSRC.rglob("*.cu"), *CSRC.rglob("*.cpp")])


def headers_digest() -> str:
    h = hashlib.sha256()
    for p in sorted([*CSRC.rglob("*.cuh"), *CSRC.rglob("*.h")]):
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def _compile_one(nvcc: str, src: Path, obj: Path, log: Path) -> tuple[Path, int, str]:
    cmd = [nvcc, *NVCC_FLAGS, "-x", "cu", "-c", str(src), "-o", str(obj)]
    p = subprocess.run(cmd, capture_output=True, text=True)
    log.write_text(" ".join(cmd) + "\n" + p.stdout + p.stderr)
    return src, p.returncode, p.stdout + p.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = find_nvcc()
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    stamp = OUT_DIR / "headers.digest"
    digest = headers_digest()
    if not stamp.exists() or stamp.read_text() != digest:
        force = True
    jobs = []
    objs = []
    for src in sources():
        rel = src.relative_to(CSRC)
        obj = OBJ_DIR / ("_".join(rel.parts) + ".o")
        objs.append(obj)
        if force or not obj.exists() or obj.stat().st_mtime < src.stat().st_mtime:
            jobs.append((src, obj, obj.with_suffix(".log")))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
            futs = [ex.submit(_compile_one, nvcc, s, o, l) for s, o, l in jobs]
            failed = []
            for f in cf.as_completed(futs):
                src, rc, out = f.result()
                if verbose or rc != 0:
                    print(f"--- {src.relative_to(ROOT)} (rc={rc})\n{out}", file=sys.stderr)
                if rc != 0:
                    failed.append(src)
            if failed:
                raise RuntimeError(f"nvcc failed for: {', '.join(str(s) for s in failed)}")
    if jobs or not LIB_PATH.exists():
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH), *map(str, objs),
               "-Xcompiler", "-fPIC", "-lpthread", "-ldl"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n" + p.stdout + p.stderr)
    stamp.write_text(digest)
    return LIB_PATH


HOST_ASAN_LIB = OUT_DIR / "libinfomesh_b200_host_asan.so"


def build_host_sanitized() -> Path:
    """The C++ host runtime alone (tokeniser / index builder / MD5 / SimHash / Hamming scan) under AddressSanitizer and
    UndefinedBehaviorSanitizer.  Load it with ``INFOMESH_B200_NATIVE_LIB=<path>`` and
    ``LD_PRELOAD=$(g++ -print-file-name=libasan.so)``; ``tests/test_native_cpu.py`` does exactly that in a subprocess."""
    srcs = sorted(CSRC.rglob("*.cpp"))
    OUT_DIR.mkdir(parents=True, exist_ok=True)
    errors = []
    for cxx in dict.fromkeys(c for c in (os.environ.get("CXX"), shutil.which("g++"), "/usr/bin/g++", shutil.which("clang++")) if c):
        cmd = [cxx, "-std=c++17", "-O1", "-g", "-fno-omit-frame-pointer", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
               "-fPIC", "-shared", "-DIM_BUILD", *map(str, srcs), "-o", str(HOST_ASAN_LIB)]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode == 0:                     # not every toolchain ships the sanitizer runtimes
            return HOST_ASAN_LIB
        errors.append(f"{cxx}: {(p.stdout + p.stderr).strip()[-300:]}")
    raise RuntimeError("sanitized host build failed:\n" + "\n".join(errors))


def sanitizer_runtime() -> str | None:
    """Path of libasan for LD_PRELOAD (python itself is not instrumented), or None when no compiler provides one."""
    for cxx in dict.fromkeys(c for c in (os.environ.get("CXX"), shutil.which("g++"), "/usr/bin/g++") if c):
        out = subprocess.run([cxx, "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
        if out and Path(out).exists():
            return str(Path(out).resolve())
    return None


def main() -> int:
    if "--sanitize-host" in sys.argv:
        print(build_host_sanitized())
        return 0
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
