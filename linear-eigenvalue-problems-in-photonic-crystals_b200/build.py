#!/usr/bin/env python3
"""Build the native library in-tree.

    python build.py            # csrc/libpcb200.so   (nvcc, sm_100a only)
    python build.py --emu      # tests/emu/_build/libpcb200_emu.so  (g++, host emulation -- TESTS ONLY)
    python build.py --sizes 48,120   # restrict the FFT plans (faster iteration)

One translation unit per grid size (csrc/pcb_plans.inc) + the C ABI; objects are cached by a
content hash of the sources so that rebuilding after a small edit is cheap.  There is no
multi-arch build: -gencode arch=compute_100a,code=sm_100a is the only target.
"""
import argparse
import concurrent.futures as cf
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
EMU_DIR = os.path.join(ROOT, "tests", "emu")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false", "-Xptxas", "-warn-spills"]
NVCC_FLAGS = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + os.environ.get("PCB200_NVCC_EXTRA", "").split()
GXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-DPCB_EMU", "-x", "c++", "-include", os.path.join(EMU_DIR, "emu_cuda.h"),
             "-I", EMU_DIR, "-Wno-unknown-pragmas", "-Wno-unused-value", "-fno-strict-aliasing"]


def plans():
    out = []
    for line in open(os.path.join(CSRC, "pcb_plans.inc")):
        m = re.match(r"\s*PCB_PLAN\((\d+),\s*(\d+),\s*(\d+)\)", line)
        if m:
            out.append(tuple(int(v) for v in m.groups()))
    return out


def src_hash(extra=""):
    h = hashlib.sha256(extra.encode())
    files = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".inc")))
    for f in files:
        h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(ROOT, "include", "pcb200.h"), "rb").read())
    if "emu" in extra:
        h.update(open(os.path.join(EMU_DIR, "emu_cuda.h"), "rb").read())
    return h.hexdigest()[:16]


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise SystemExit(f"build failed: {cmd[0]}")
    if r.stderr.strip():
        sys.stderr.write(r.stderr)


def build(emu=False, sizes=None, jobs=None, verbose=True):
    sel = plans()
    if sizes:
        sel = [p for p in sel if p[0] in sizes]
    tag = ("emu" if emu else "cuda") + ",".join(str(p[0]) for p in sel)
    bdir = os.path.join(EMU_DIR, "_build") if emu else os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    lib = os.path.join(bdir, "libpcb200_emu.so") if emu else os.path.join(CSRC, "libpcb200.so")
    stamp = os.path.join(bdir, ("emu" if emu else "cuda") + ".stamp")
    h = src_hash(tag)
    if os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == h:
        if verbose:
            print(f"[build] {os.path.relpath(lib, ROOT)} up to date")
        return lib
    # a restricted plan list is written next to the objects so the C ABI only references built plans
    inc_dir = bdir
    with open(os.path.join(inc_dir, "pcb_plans.inc"), "w") as f:
        for n, r1, r2 in sel:
            f.write(f"PCB_PLAN({n}, {r1}, {r2})\n")
    cc = ["g++"] + GXX_FLAGS if emu else [NVCC] + NVCC_FLAGS
    jobs_list = []
    objs = []
    for n, r1, r2 in sel:
        o = os.path.join(bdir, f"op_{n}.o")
        objs.append(o)
        jobs_list.append(cc + [f"-DPCB_N={n}", f"-DPCB_R1={r1}", f"-DPCB_R2={r2}", "-I", CSRC, "-c",
                               os.path.join(CSRC, "pcb_operator_inst.cu"), "-o", o])
    o = os.path.join(bdir, "capi.o")
    objs.append(o)
    # -I bdir first: picks up the restricted pcb_plans.inc
    jobs_list.append(cc + ["-I", inc_dir, "-I", CSRC, "-DPCB_PLANS_FROM_BUILD", "-c", os.path.join(CSRC, "pcb_capi.cu"), "-o", o])
    with cf.ThreadPoolExecutor(max_workers=jobs or os.cpu_count() or 4) as ex:
        list(ex.map(run, jobs_list))
    if emu:
        run(["g++", "-shared", "-o", lib] + objs)
    else:
        run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs)
    with open(stamp, "w") as f:
        f.write(h)
    if verbose:
        print(f"[build] wrote {os.path.relpath(lib, ROOT)} ({len(sel)} grid sizes)")
    return lib


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--emu", action="store_true")
    ap.add_argument("--sizes", default="")
    ap.add_argument("-j", type=int, default=None)
    a = ap.parse_args()
    build(emu=a.emu, sizes=[int(s) for s in a.sizes.split(",") if s] or None, jobs=a.j)
