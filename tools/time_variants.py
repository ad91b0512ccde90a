"""A/B of library builds on one box: python tools/time_variants.py N lattice type cols lib1.so lib2.so ...  (one subprocess per library)."""
import json
import os
import subprocess
import sys

N, lat, typ, cols = sys.argv[1:5]
here = os.path.dirname(os.path.abspath(__file__))
code = """
import importlib, os, sys, json, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(%r))
os.environ.setdefault("PCB200_QUIET", "1")
pcb = importlib.import_module("linear-eigenvalue-problems-in-photonic-crystals_b200")
pcb._lib.use_library(sys.argv[1])
N, d_flag, typ, m = int(sys.argv[2]), sys.argv[3], sys.argv[4], int(sys.argv[5])
mfd, ne, L = pcb.discretization, pcb.numerical_experiments, pcb._lib
alpha = pcb.dielectric.kpath(d_flag)[0]
relax, pnt = mfd.set_relaxation(alpha)
a_fft, b_fft = mfd.fft_blocks(N, 1, pcb.dielectric.diel_info(d_flag, option="ct"), alpha=alpha)
inv_fft = mfd.inverse_3_times_3_B(b_fft, pnt, relax[0])
Diels = None if typ == "none" else getattr(mfd, typ + "_handle")(N, d_flag)
A, H, P = ne.pc_mfd_handle(a_fft, (pnt * b_fft[0], pnt * b_fft[1]), Diels, inv_fft, relax[0])
ctx = pcb.get_context(N)
X, Y = ctx.random_block(m, 1), ctx.empty(m)
for _ in range(3): H.op.apply_into(L.APPLY_H, X, Y)
ctx.sync(); ctx.timer_start()
for _ in range(20): H.op.apply_into(L.APPLY_H, X, Y)
total = ctx.timer_stop() / 20
buf, npass = (C.c_float * 8)(), C.c_int()
acc = np.zeros(8)
for _ in range(5):
    L.check(L.lib().pcb_apply_timed(H.op.h, L.APPLY_H, m, L.ptr_array(X.ptrs), L.ptr_array(Y.ptrs), buf, C.byref(npass)), "timed")
    acc += np.array(buf[:8])
print(json.dumps({"lib": os.path.basename(sys.argv[1]), "ms": round(total, 4), "pass_ms": [round(float(v) / 5, 4) for v in acc[:npass.value]]}))
""" % here
for lib in sys.argv[5:]:
    r = subprocess.run([sys.executable, "-c", code, os.path.abspath(lib), N, lat, typ, cols], capture_output=True, text=True)
    print((r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1], flush=True)
