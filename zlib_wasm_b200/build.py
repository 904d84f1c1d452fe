"""Build everything in-tree.

  zlib_wasm_b200/libzb200.so   the product: CUDA kernels (sm_100a) + C ABI + zlib.h-compatible host API
  tools/libzgen.so             synthetic data generator (tests / bench)
  oracle/liboracle.so          CPU restatement          (tests / bench cpu_baseline only)
  oracle/_ref/*                the unmodified reference, only where /root/reference exists

nvcc cross-compiles for sm_100a without a GPU.  Outputs are rebuilt only when a
source is newer, so calling build_all() from every test session is cheap.
"""
import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libzb200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "--expt-extended-lambda",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unknown-pragmas",
              "-Xptxas", "-v"]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-Wall", "-Wno-unknown-pragmas"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, log=None):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    return p.stdout


def _cuda_home():
    for c in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if c and os.path.exists(os.path.join(c, "bin", "nvcc")):
            return c
    nv = shutil.which("nvcc")
    if nv:
        return os.path.dirname(os.path.dirname(nv))
    raise RuntimeError("nvcc not found")


def build_product(verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    cuda = _cuda_home()
    nvcc = os.path.join(cuda, "bin", "nvcc")
    headers = glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h"))
    objs = []
    procs = []
    for src in sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp"))):
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        objs.append(obj)
        if not _newer(obj, [src] + headers):
            continue
        if src.endswith(".cu"):
            cmd = [nvcc] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj]
        else:
            cmd = ["g++"] + CXX_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
                                         "-c", src, "-o", obj]
        procs.append((src, obj, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, obj, cmd, p in procs:
        out, _ = p.communicate()
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("compile failed: " + src)
        if verbose:
            sys.stderr.write(out)
    vmap = os.path.join(CSRC, "libzb200.map")                   # only the C ABI + the zlib / wasm_module names are exported
    if procs or _newer(LIB, objs + [vmap]):
        _run([nvcc, "-shared", "-o", LIB] + objs +
             ["-cudart", "static", "-Xlinker", "-Bsymbolic", "-Xlinker", "--exclude-libs,ALL",
              "-Xlinker", "--version-script=" + vmap, "-lpthread", "-ldl", "-lrt"])
    return LIB


def build_tools():
    src = os.path.join(ROOT, "tools", "zgen.c")
    out = os.path.join(ROOT, "tools", "libzgen.so")
    if _newer(out, [src]):
        _run(["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-o", out, src])
    return out


def build_oracle():
    od = os.path.join(ROOT, "oracle")
    out = os.path.join(od, "liboracle.so")
    if _newer(out, [os.path.join(od, "zoracle.c"), os.path.join(od, "zoracle.h")]):
        _run(["make", "-C", od, "liboracle.so"])
    if os.path.exists("/root/reference/deflate.c"):             # make decides what is out of date
        _run(["make", "-C", od, "ref"])
    return out


def build_dropin_demo():
    """The reference's own examples/zpipe.c (compiled against the reference's
    zlib.h, untouched) linked to libzb200.so: the drop-in boundary exercised by
    a reference caller.  Only possible where /root/reference exists; the binary
    is git-ignored and travels with the gpurun snapshot."""
    src = "/root/reference/examples/zpipe.c"
    out = os.path.join(ROOT, "tests", "_bin", "zpipe_b200")
    if not os.path.exists(src):
        return None
    if _newer(out, [src, LIB]):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        _run(["gcc", "-O2", "-w", "-I/root/reference", "-o", out, src, "-L" + PKG, "-lzb200",
              "-Wl,-rpath,$ORIGIN/../../zlib_wasm_b200"])
    # examples/gun.c: gunzip through inflateBack() (infback.c) + crc32()
    src2, out2 = "/root/reference/examples/gun.c", os.path.join(ROOT, "tests", "_bin", "gun_b200")
    if os.path.exists(src2) and _newer(out2, [src2, LIB]):
        _run(["gcc", "-O2", "-w", "-I/root/reference", "-o", out2, src2, "-L" + PKG, "-lzb200",
              "-Wl,-rpath,$ORIGIN/../../zlib_wasm_b200"])
    # examples/zran.c (random access: inflate(Z_BLOCK) + data_type, inflatePrime, inflateSetDictionary), gzjoin.c and
    # gzappend.c (Z_BLOCK, deflatePrime): the reference's own users of the block-level API, untouched
    for name, extra in (("zran", ["-DTEST"]), ("gzjoin", []), ("gzappend", [])):
        srcn = "/root/reference/examples/%s.c" % name
        outn = os.path.join(ROOT, "tests", "_bin", name + "_b200")
        if os.path.exists(srcn) and _newer(outn, [srcn, LIB]):
            _run(["gcc", "-O2", "-w", "-I/root/reference", "-I/root/reference/examples"] + extra + ["-o", outn, srcn, "-L" + PKG, "-lzb200",
                  "-Wl,-rpath,$ORIGIN/../../zlib_wasm_b200"])
    return out


def build_all(verbose=False):
    build_tools()
    build_oracle()
    lib = build_product(verbose)
    build_dropin_demo()
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import zlib_wasm_b200  # noqa: F401  (import check)
    return lib


if __name__ == "__main__":
    print(build_all(verbose="-v" in sys.argv))
