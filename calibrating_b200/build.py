"""In-tree build of libb2s.so (hand-written sm_100a CUDA + the C-ABI of include/b2s.h).

    python -m calibrating_b200.build        # nvcc cross-compiles without a GPU

Every .cu is compiled to an object file in csrc/_obj/ (in parallel, only when it or a header changed), then linked.
The .so is git-ignored but travels to the GPU box with the gpurun snapshot; its build hash (sha256 over the sources) is
exported as b2s_build_hash() so that profiles can be tied to the binary they were measured on.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "_obj")
OUT = os.path.join(PKG, "libb2s.so")
SOURCES = ["b2s_api.cu", "sgbm_cost.cu", "sgbm_agg.cu", "sgbm_wave.cu", "sgbm_sweep6.cu", "sgbm_post.cu", "remap.cu", "resize.cu", "cloud.cu"]
HEADERS = [os.path.join(CSRC, "b2s_internal.h"), os.path.join(CSRC, "sgm_common.cuh"), os.path.join(ROOT, "include", "b2s.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def source_hash():
    """sha256 over the CUDA sources and headers (first 16 hex digits): identifies the build a profile belongs to."""
    h = hashlib.sha256()
    for f in sorted([os.path.join(CSRC, s) for s in SOURCES] + HEADERS):
        if os.path.exists(f):
            h.update(os.path.basename(f).encode())
            h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def needs_build():
    return _stale(OUT, [os.path.join(CSRC, s) for s in SOURCES] + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hash_hdr = os.path.join(OBJ, "build_hash.h")
    text = '#define B2S_BUILD_HASH "%s"\n' % source_hash()
    if not os.path.exists(hash_hdr) or open(hash_hdr).read() != text:
        open(hash_hdr, "w").write(text)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(s):
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        deps = [src] + HEADERS + ([hash_hdr] if s == "b2s_api.cu" else [])
        if force or _stale(obj, deps):
            cmd = [nvcc] + NVCC_FLAGS + ["-I", OBJ] + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
            subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", OUT] + objs)
    return OUT


if __name__ == "__main__":
    print(build(force="-f" in sys.argv or "-v" in sys.argv, verbose="-v" in sys.argv))
