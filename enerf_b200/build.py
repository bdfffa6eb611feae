"""Builds enerf_b200/libenerf_b200.so in-tree with nvcc for sm_100a (no torch headers involved: the
library is a plain C ABI, see include/enerf_b200.h).

    python -m enerf_b200.build [--force]

Objects are cached under enerf_b200/csrc/build/ keyed by a hash of (source, headers, flags); the
translation units compile in parallel (the fused ray kernel is instantiated once per object file).
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libenerf_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]

# (object name, source, extra -D flags)
RAY_INSTANCES = [(11, 2, 1), (11, 3, 1), (11, 4, 1), (11, 8, 0), (35, 2, 1), (35, 3, 1), (35, 4, 1), (35, 8, 0)]


def units():
    u = [(n, n + ".cu", []) for n in ("camera", "feature_net", "cost_volume", "cost_reg", "render_rays", "render_rays_tc", "render_rays_ws", "tc_selftest", "tc_probe", "tc_conv", "tc_conv2", "mask_rays", "eval_ops", "composite")]
    for fc, s, st in RAY_INSTANCES:
        u.append((f"rr_{fc}_{s}_{st}", "render_rays_inst.cu", [f"-DRR_FC={fc}", f"-DRR_S={s}", f"-DRR_STATIC={st}"]))
    for sv in range(2, 9):
        u.append((f"rtc_{sv}", "render_rays_tc_inst.cu", [f"-DRTC_S={sv}"]))
    extra = os.path.join(CSRC, "units.txt")  # optional additional units: "<name> <source> [-Dflags...]"
    if os.path.exists(extra):
        for line in open(extra):
            parts = line.split()
            if parts and not parts[0].startswith("#"):
                u.append((parts[0], parts[1], parts[2:]))
    return u


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libenerf_b200.so")


def _headers_digest():
    h = hashlib.sha256()
    for d in (CSRC, INCLUDE):
        for f in sorted(os.listdir(d)):
            if f.endswith((".cuh", ".h")):
                h.update(f.encode())
                h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()


def _compile(nvcc, name, src, defs, hdr_digest, force):
    obj = os.path.join(OBJ, name + ".o")
    key = hashlib.sha256((hdr_digest + " ".join(NVCC_FLAGS + defs)).encode() + open(os.path.join(CSRC, src), "rb").read()).hexdigest()
    stamp = obj + ".key"
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == key:
        return obj, False
    cmd = [nvcc] + NVCC_FLAGS + defs + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    open(stamp, "w").write(key)
    return obj, True


def build(force=False, verbose=True):
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdr = _headers_digest()
    us = units()
    objs, rebuilt = [], 0
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(len(us), os.cpu_count() or 4))) as ex:
        futs = [ex.submit(_compile, nvcc, n, s, d, hdr, force) for n, s, d in us]
        for f in futs:
            obj, did = f.result()
            objs.append(obj)
            rebuilt += did
    if rebuilt or not os.path.exists(LIB):
        # visibility=hidden for C++ symbols; the extern "C" entry points are exported explicitly
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"[enerf_b200.build] {rebuilt} unit(s) compiled, library: {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
