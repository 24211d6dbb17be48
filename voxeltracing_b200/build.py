"""In-tree native builds: the CUDA product library, the CPU oracle (test infrastructure) and, when the
reference is mounted, the reference-shader build under oracle/_ref (test infrastructure).

Everything is built with explicit compiler invocations so the resulting .so files live in the tree
and travel to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "voxeltracing_b200" / "csrc"
CUDA_LIB = ROOT / "voxeltracing_b200" / "libvxrt_cuda.so"
HOST_LIB = ROOT / "voxeltracing_b200" / "libvxrt_host.so"
ORACLE_LIB = ROOT / "oracle" / "libvxrt_oracle.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",            # parity: no FMA contraction anywhere on the ray path
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared",
]
GXX_ORACLE_FLAGS = ["-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths, extra="") -> str:
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        h.update(str(p).encode())
        h.update(Path(p).read_bytes())
    return h.hexdigest()


def _up_to_date(target: Path, stamp: str) -> bool:
    s = target.with_suffix(target.suffix + ".stamp")
    return target.exists() and s.exists() and s.read_text() == stamp


def _write_stamp(target: Path, stamp: str) -> None:
    target.with_suffix(target.suffix + ".stamp").write_text(stamp)


def _run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(map(str, cmd)), r.stdout))
    return r.stdout


def build_cuda_variant(name: str, defines, verbose: bool = False) -> Path:
    """Experimental build with extra -D flags into voxeltracing_b200/libvxrt_cuda_<name>.so (select it with VXRT_CUDA_LIB)."""
    out = CUDA_LIB.with_name(f"libvxrt_cuda_{name}.so")
    cmd = [_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(out)] + [str(s) for s in sorted(CSRC.glob("*.cu"))]
    o = _run(cmd)
    if verbose:
        print(o)
    return out


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    srcs = sorted(CSRC.glob("*.cu"))
    deps = srcs + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [ROOT / "include" / "vxrt_cuda.h"]
    stamp = _digest(deps, " ".join(NVCC_FLAGS))
    if not force and _up_to_date(CUDA_LIB, stamp):
        return CUDA_LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(CUDA_LIB)] + [str(s) for s in srcs]
    out = _run(cmd)
    if verbose:
        print(out)
    _write_stamp(CUDA_LIB, stamp)
    return CUDA_LIB


def build_host(force: bool = False) -> Path:
    hdir = ROOT / "voxeltracing_b200" / "host"
    srcs = sorted(hdir.glob("*.cpp"))
    if not srcs:
        return HOST_LIB
    deps = srcs + sorted(hdir.glob("*.h"))
    flags = ["-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared"]
    stamp = _digest(deps, " ".join(flags))
    if not force and _up_to_date(HOST_LIB, stamp):
        return HOST_LIB
    _run(["g++"] + flags + ["-o", str(HOST_LIB)] + [str(s) for s in srcs] + ["-lz"])  # zlib: region-file chunks (vxrt_mca.cpp)
    _write_stamp(HOST_LIB, stamp)
    return HOST_LIB


def build_oracle(force: bool = False) -> Path:
    odir = ROOT / "oracle"
    # ref_driver.cpp belongs to oracle/_ref (it needs files build_ref.py generates), not to the restatement
    srcs = sorted(p for p in odir.glob("*.cpp") if p.name != "ref_driver.cpp")
    deps = srcs + sorted(p for p in odir.glob("*.h") if p.name != "glsl_shim.h") + [ROOT / "include" / "vxrt_cuda.h"]
    stamp = _digest(deps, " ".join(GXX_ORACLE_FLAGS))
    if not force and _up_to_date(ORACLE_LIB, stamp):
        return ORACLE_LIB
    _run(["g++"] + GXX_ORACLE_FLAGS + ["-o", str(ORACLE_LIB)] + [str(s) for s in srcs])
    _write_stamp(ORACLE_LIB, stamp)
    return ORACLE_LIB


def build_ref(force: bool = False):
    """oracle/_ref: the reference's own shader sources compiled for the CPU (only where the reference
    tree is mounted; the prebuilt .so travels to the GPU box)."""
    script = ROOT / "oracle" / "build_ref.py"
    if not script.exists() or not Path(os.environ.get("VXRT_REFERENCE", "/root/reference")).exists():
        return None
    _run([sys.executable, str(script)] + (["--force"] if force else []))
    world = ROOT / "oracle" / "build_ref_world.py"   # world producers: FastNoise, enkiMI, WorldGenerator.cpp, Importer.cpp
    if world.exists():
        _run([sys.executable, str(world)] + (["--force"] if force else []))
    return ROOT / "oracle" / "_ref" / "libvxrt_ref.so"


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda(force, verbose)
    build_host(force)
    build_oracle(force)
    build_ref(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", CUDA_LIB, ORACLE_LIB)
