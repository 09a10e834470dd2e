"""In-tree build of libkmc.so (CUDA, sm_100a only) and the kmer-count CLI.  nvcc cross-compiles
without a GPU; the built files are git-ignored but travel to the GPU box with the snapshot."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def lib_path():
    return os.environ.get("KMC_LIB") or os.path.join(_HERE, "libkmc.so")


def cli_path():
    return os.path.join(_HERE, "bin", "kmer-count")


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources():
    src = [os.path.join(_CSRC, f) for f in sorted(os.listdir(_CSRC)) if f.endswith((".cu", ".cuh", ".cpp", ".hpp", ".h"))]
    src.append(os.path.join(os.path.dirname(_HERE), "include", "kmc.h"))
    return src


def build(force=False, verbose=False):
    """Compile libkmc.so and bin/kmer-count if sources changed.  Returns the library path.

    Safe to call from several processes at once (one rank per GPU under torchrun): one of them compiles, under a
    file lock, into a temporary file that is renamed into place; the others wait and find the result up to date."""
    srcs = _sources()
    lib = lib_path()
    if os.environ.get("KMC_LIB"):
        return lib  # an explicitly chosen library (experiments): use it as it is
    cli = cli_path()
    cli_src = os.path.join(_CSRC, "kmc_cli.cpp")
    if not force and _newer(lib, srcs) and (not os.path.exists(cli_src) or _newer(cli, srcs + [lib])):
        return lib
    import fcntl
    with open(os.path.join(_HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if force or not _newer(lib, srcs):  # checked again: another process may have built it while we waited
            tmp = f"{lib}.{os.getpid()}.tmp"
            cmd = [NVCC, "-O3", "-std=c++17", *ARCH, "-lineinfo", "-Xcompiler", "-fPIC", "-shared",
                   "-o", tmp, os.path.join(_CSRC, "kmc_api.cu")]
            if verbose:
                cmd += ["-Xptxas", "-v"]
            try:
                subprocess.run(cmd, check=True)
                os.replace(tmp, lib)
            finally:
                if os.path.exists(tmp):
                    os.remove(tmp)
        if os.path.exists(cli_src) and (force or not _newer(cli, srcs + [lib])):
            os.makedirs(os.path.dirname(cli), exist_ok=True)
            tmp = f"{cli}.{os.getpid()}.tmp"
            try:
                subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", tmp, cli_src, "-L" + _HERE, "-lkmc",
                                "-Wl,-rpath,$ORIGIN/.."], check=True)
                os.replace(tmp, cli)
            finally:
                if os.path.exists(tmp):
                    os.remove(tmp)
    return lib
