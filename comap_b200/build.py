"""Builds libcomap_b200.so (CUDA sm_100a + C ABI) in-tree with nvcc.

    python -m comap_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libcomap_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CU = ["capi.cu", "capi_stats.cu", "capi_inter.cu", "capi_mica.cu", "k5_mica.cu", "k1_map.cu", "k1_mma.cu", "k1_mma20.cu", "k1_variants.cu", "k2_pairs.cu", "k3_simulate.cu", "k4_cluster.cu", "k4_rnn.cu"]
CPP = ["tables.cpp", "schedule.cpp", "comm.cpp"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "comap_b200.h"))
    objs = []
    procs = []
    for f in CU + CPP:
        src = os.path.join(CSRC, f)
        obj = os.path.join(objdir, f + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [NVCC] + ARCH + COMMON + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu" if f.endswith(".cu") else "c++", "-c", src, "-o", obj]
            procs.append((f, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for f, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s\n%s\n" % (f, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(OUT, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", OUT] + objs + ["-lcudart", "-ldl"]
        subprocess.check_call(cmd)
    return OUT


HOST_SRC = ["main.cpp", "options.cpp", "seq.cpp", "tree.cpp", "models.cpp"]
HOST_BIN = os.path.join(HERE, "bin", "comap_b200")


def build_host(force=False):
    """Builds the C++ front-end (CoMap's command line over the C ABI) with g++."""
    hdir = os.path.join(HERE, "host")
    srcs = [os.path.join(hdir, f) for f in HOST_SRC]
    deps = srcs + [os.path.join(hdir, "bpp.h"), os.path.join(os.path.dirname(HERE), "include", "comap_b200.h"), OUT]
    os.makedirs(os.path.dirname(HOST_BIN), exist_ok=True)
    if force or _stale(HOST_BIN, deps):
        cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-Wall", "-Wextra", "-pthread", "-o", HOST_BIN] + srcs + \
              ["-L" + HERE, "-lcomap_b200", "-Wl,-rpath,$ORIGIN/..", "-Wl,-rpath-link," + "/usr/local/cuda/lib64"]
        subprocess.check_call(cmd)
    # `mica` (CoMap/Mica.cpp) is the same executable under its own name (main() dispatches on argv[0])
    mica = os.path.join(os.path.dirname(HOST_BIN), "mica_b200")
    if force or _stale(mica, [HOST_BIN]):
        import shutil
        shutil.copy2(HOST_BIN, mica)
    return HOST_BIN


if __name__ == "__main__":
    print(build_host(force="--force" in sys.argv) if "--host" in sys.argv else build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
