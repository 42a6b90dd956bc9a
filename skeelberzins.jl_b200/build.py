"""Build libskeelberzins_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

One translation unit per registry functor (csrc/sb_inst.cu with -DSB_FN=...) plus the C-ABI unit
(csrc/sb_api.cu), compiled in parallel and linked into skeelberzins.jl_b200/libskeelberzins_b200.so.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# tuning builds: SB_EXTRA_NVCC_FLAGS="-DSB_TMA_STAGES=2" SB_LIB_SUFFIX=_s2 python build.py  -> libskeelberzins_b200_s2.so
_SUFFIX = os.environ.get("SB_LIB_SUFFIX", "")
OBJ = os.path.join(HERE, "build" + _SUFFIX)
LIB = os.path.join(HERE, "libskeelberzins_b200%s.so" % _SUFFIX)

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC"] + ARCH + os.environ.get("SB_EXTRA_NVCC_FLAGS", "").split()

FUNCTORS = [
    ("FnLinDiff", "entry_lin_diff", "LIN_DIFF"),
    ("FnNonlinDiff", "entry_nonlin_diff", "NONLIN_DIFF"),
    ("FnLinDiffCyl", "entry_lin_diff_cyl", "LIN_DIFF_CYL"),
    ("FnPoisson", "entry_poisson", "POISSON"),
    ("FnStatNonlin", "entry_stat_nonlin", "STAT_NONLIN"),
    ("FnReactDiff", "entry_react_diff", "REACT_DIFF"),
    ("FnLinDiffSph", "entry_lin_diff_sph", "LIN_DIFF_SPH"),
    ("FnConvDiff", "entry_conv_diff", "CONV_DIFF"),
    ("FnLinReact", "entry_lin_react", "LIN_REACT"),
    ("FnGenericScalar", "entry_generic_scalar", "GENERIC_SCALAR"),
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources_digest():
    return _sources_digest()


def _sources_digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def build(force=False, verbose=False, ptxas_verbose=False):
    """Compile everything; returns the path of the shared library."""
    digest = _sources_digest()
    stamp = os.path.join(OBJ, "stamp")
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    flags = list(NVCC_FLAGS)
    if ptxas_verbose:
        flags += ["-Xptxas", "-v"]
    jobs = []
    for fn, entry, name in FUNCTORS:
        obj = os.path.join(OBJ, "inst_%s.o" % name.lower())
        jobs.append((obj, [nvcc] + flags + ["-DSB_FN=%s" % fn, "-DSB_ENTRY=%s" % entry, '-DSB_NAME="%s"' % name,
                                            "-c", os.path.join(CSRC, "sb_inst.cu"), "-o", obj]))
    api_obj = os.path.join(OBJ, "sb_api.o")
    jobs.append((api_obj, [nvcc] + flags + ['-DSB_SOURCE_DIGEST="%s"' % digest, "-c", os.path.join(CSRC, "sb_api.cu"), "-o", api_obj]))
    logs = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(8, os.cpu_count() or 1))) as ex:
        for out in ex.map(lambda j: _run(j[1]), jobs):
            logs.append(out)
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(nvcc)), "lib64")
    link = [nvcc, "-shared", "-o", LIB] + ARCH + [j[0] for j in jobs] + ["-lcudart", "-lnvrtc", "-ldl", "-Xlinker", "-rpath=" + cuda_lib]
    logs.append(_run(link))
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose or ptxas_verbose:
        sys.stdout.write("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, ptxas_verbose="--ptxas" in sys.argv))
