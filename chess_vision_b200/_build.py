"""Builds libchessvision_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libchessvision_b200.so")
SOURCES = ["api.cu", "kernels_generic.cu", "kernels_umma.cu", "kernels_exact.cu", "kernels_frontend.cu", "kernels_frontend3.cu", "kernels_backend.cu", "kernels_head.cu", "fen.cu", "synth.cu", "eval.cu", "resize.cu", "pack.cu", "jpeg.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "--expt-relaxed-constexpr"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libchessvision_b200.so cannot be built (there is no CPU fallback)")


def _src_hash():
    """Content hash of everything the library is compiled from (mtimes do not survive the copy to the GPU box)."""
    import hashlib
    hsh = hashlib.sha1()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(HERE, "..", "include", "chessvision_b200.h")]
    for f in files:
        hsh.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            hsh.update(fh.read())
    hsh.update(" ".join(NVCC_FLAGS + SOURCES).encode())
    return hsh.hexdigest()


def _stale():
    """True when the library is missing or was built from other sources (hash stored beside it by _build)."""
    if not os.path.exists(LIB):
        return True
    try:
        with open(LIB + ".srchash") as fh:
            return fh.read().strip() != _src_hash()
    except OSError:
        return True


def build(force=False, verbose=False, variant=None, extra=()):
    """Compile every .cu under csrc/ into one shared library.  Returns the library path.
    variant="prof", extra=["-DCV_FE_PROFILE"] builds libchessvision_b200_prof.so beside it (experiments; see CV_B200_LIB)."""
    global LIB
    if variant:
        saved = LIB
        LIB = os.path.join(HERE, f"libchessvision_b200_{variant}.so")
        try:
            return _build(True, verbose, os.path.join(HERE, "build_" + variant), list(extra))
        finally:
            LIB = saved
    if not force and not _stale():
        return LIB
    return _build(force, verbose, os.path.join(HERE, "build"), [])


def _build(force, verbose, build_dir, extra):
    objs = []
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, *os.environ.get("CV_NVCC_EXTRA", "").split(), *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
    link = [_nvcc(), "-shared", "-o", LIB + ".tmp", *objs, "-gencode", "arch=compute_100a,code=sm_100a",
            ]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(LIB + ".tmp", LIB)
    if not extra and not os.environ.get("CV_NVCC_EXTRA"):
        with open(LIB + ".srchash", "w") as fh:
            fh.write(_src_hash())
    return LIB


if __name__ == "__main__":
    import sys
    var = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=var[0] if var else None,
                extra=[a for a in sys.argv[1:] if a.startswith("-D")]))
