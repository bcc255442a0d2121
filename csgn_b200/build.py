"""In-tree build of the native code (nvcc cross-compiles sm_100a without a GPU).

  csgn_b200/lib/libcsgn.so      CUDA kernels + C ABI (include/csgn.h)
  csgn_b200/lib/libcertFHE.so   the certFHE C++ drop-in classes over that C ABI
  tests/cpp/bin/*               C++ acceptance / differential executables

Everything is written for sm_100a only.  Built files are git-ignored but travel to
the GPU box with the gpurun snapshot.
"""
import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
CERTFHE = os.path.join(PKG, "certfhe")
LIBDIR = os.path.join(PKG, "lib")
INCLUDE = os.path.join(ROOT, "include")
CPP_TESTS = os.path.join(ROOT, "tests", "cpp")

NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    if verbose and r.stdout.strip():
        print(r.stdout)


def libcsgn_path():
    return os.path.join(LIBDIR, "libcsgn.so")


def libcertfhe_path():
    return os.path.join(LIBDIR, "libcertFHE.so")


def build_libcsgn(force=False, verbose=False, variants=False):
    """libcsgn.so: one kernel per shape.  variants=True builds libcsgn_variants.so instead, with the losing kernel
    variants of the tuning sweeps compiled in (-DCSGN_BUILD_VARIANTS; select it with CSGN_LIBRARY=<path> for A/B runs)."""
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    out = os.path.join(LIBDIR, "libcsgn_variants.so") if variants else libcsgn_path()
    if force or _stale(out, deps):
        extra = ["-DCSGN_BUILD_VARIANTS"] if variants else []
        _run([_nvcc()] + NVCC_ARCH + NVCC_FLAGS + extra + ["-shared", "-o", out] + srcs, verbose)
    return out


def build_libcertfhe(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(CERTFHE, "*.cpp")))
    if not srcs:
        return None
    deps = srcs + glob.glob(os.path.join(CERTFHE, "*.h")) + glob.glob(os.path.join(INCLUDE, "*.h")) + [libcsgn_path()]
    out = libcertfhe_path()
    if force or _stale(out, deps):
        # -O3: the Bitlen validation of every Ciphertext(V, Bitlen, len, ctx) is one long loop that -O3 vectorises
        _run(["g++", "-O3", "-std=c++11", "-fPIC", "-shared", "-Wall", "-I" + INCLUDE, "-I" + CERTFHE, "-o", out]
             + srcs + ["-L" + LIBDIR, "-lcsgn", "-Wl,-rpath,$ORIGIN"], verbose)
    return out


REFERENCE = os.environ.get("CSGN_REFERENCE", "/root/reference")


def build_cpp_tests(force=False, verbose=False):
    """C++ executables under tests/cpp, written against the certFHE class API.

    diff_vs_reference.cpp additionally needs the reference's HEADERS at compile time and
    links oracle/_ref/libcertfhe_ref.so; it is (re)built only where /root/reference is
    present -- the binary travels to the GPU box."""
    outs = []
    bindir = os.path.join(CPP_TESTS, "bin")
    srcs = sorted(glob.glob(os.path.join(CPP_TESTS, "*.cpp")))
    if not srcs or not os.path.exists(libcertfhe_path()):
        return outs
    os.makedirs(bindir, exist_ok=True)
    hdrs = glob.glob(os.path.join(CERTFHE, "*.h"))
    rpath = "-Wl,-rpath,$ORIGIN/../../../csgn_b200/lib"
    for src in srcs:
        name = os.path.splitext(os.path.basename(src))[0]
        out = os.path.join(bindir, name)
        cmd = ["g++", "-O2", "-std=c++11", "-Wall", "-I" + CERTFHE, "-I" + INCLUDE, "-o", out, src,
               "-L" + LIBDIR, "-lcertFHE", "-lcsgn", rpath]
        deps = [src, libcertfhe_path()] + hdrs
        if name in ("diff_vs_reference", "host_vs_reference"):
            ref_hdr = os.path.join(REFERENCE, "src", "certFHE.h")
            ref_lib = os.path.join(ROOT, "oracle", "_ref", "libcertfhe_ref.so")
            if not (os.path.exists(ref_hdr) and os.path.exists(ref_lib)):
                if os.path.exists(out):
                    outs.append(out)
                continue
            cmd += ["-w", '-DCSGN_REFERENCE_HEADER="%s"' % ref_hdr, "-L" + os.path.dirname(ref_lib), "-lcertfhe_ref",
                    "-pthread", "-Wl,-rpath,$ORIGIN/../../../oracle/_ref"]
            deps.append(ref_lib)
        if force or _stale(out, deps):
            _run(cmd, verbose)
        outs.append(out)
    return outs


def build_reference_demos(force=False, verbose=False):
    """Source-level drop-in check: the reference's OWN demo programs (tests/*.cpp, which say
    `#include "../src/certFHE.h"`) compiled UNMODIFIED against this repository's headers.

    Nothing is copied: build/dropin/tests/*.cpp are symlinks into /root/reference and
    build/dropin/src is a symlink to csgn_b200/certfhe, so the relative include lands on our
    certFHE.h.  Only possible where /root/reference exists; the binaries travel."""
    ref_tests = sorted(glob.glob(os.path.join(REFERENCE, "tests", "*.cpp")))
    outs = []
    if not ref_tests or not os.path.exists(libcertfhe_path()):
        return outs
    base = os.path.join(ROOT, "build", "dropin")
    os.makedirs(os.path.join(base, "tests"), exist_ok=True)
    os.makedirs(os.path.join(base, "bin"), exist_ok=True)
    link = os.path.join(base, "src")
    if not os.path.islink(link):
        os.symlink(CERTFHE, link)
    for src in ref_tests:
        name = os.path.basename(src)
        sl = os.path.join(base, "tests", name)
        if not os.path.islink(sl):
            os.symlink(src, sl)
        out = os.path.join(base, "bin", "tester_" + os.path.splitext(name)[0])
        if force or _stale(out, [src, libcertfhe_path()] + glob.glob(os.path.join(CERTFHE, "*.h"))):
            _run(["g++", "-O2", "-std=c++11", "-w", "-I" + INCLUDE, "-o", out, sl, "-L" + LIBDIR, "-lcertFHE", "-lcsgn",
                  "-Wl,-rpath,$ORIGIN/../../../csgn_b200/lib"], verbose)
        outs.append(out)
    return outs


def build_tools(force=False, verbose=False):
    """tools/bin/cpp_e2e: the bench step through the C++ drop-in API (bench.py reports it as e2e_cpp);
    tools/bin/ctor_probe: where the host time of the Ciphertext constructor goes."""
    if not os.path.exists(libcertfhe_path()):
        return []
    bindir = os.path.join(ROOT, "tools", "bin")
    os.makedirs(bindir, exist_ok=True)
    outs = []
    for name in ("cpp_e2e", "ctor_probe"):
        src = os.path.join(ROOT, "tools", name + ".cpp")
        if not os.path.exists(src):
            continue
        out = os.path.join(bindir, name)
        if force or _stale(out, [src, libcertfhe_path()] + glob.glob(os.path.join(CERTFHE, "*.h"))):
            _run(["g++", "-O2", "-std=c++11", "-Wall", "-I" + CERTFHE, "-I" + INCLUDE, "-o", out, src, "-L" + LIBDIR,
                  "-lcertFHE", "-lcsgn", "-Wl,-rpath,$ORIGIN/../../csgn_b200/lib"], verbose)
        outs.append(out)
    return outs


def build_all(force=False, verbose=False):
    build_libcsgn(force, verbose)
    build_libcertfhe(force, verbose)
    build_cpp_tests(force, verbose)
    build_reference_demos(force, verbose)
    build_tools(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
    print("built:", libcsgn_path())
    if "--variants" in sys.argv:
        print("built:", build_libcsgn(force="--force" in sys.argv, verbose=True, variants=True))
