"""Per-pass device times of one block apply: python tools/time_apply.py N d_flag type [cols]"""
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PCB200_QUIET", "1")
pcb = importlib.import_module("linear-eigenvalue-problems-in-photonic-crystals_b200")
N, d_flag, typ = int(sys.argv[1]), sys.argv[2], sys.argv[3]
m = int(sys.argv[4]) if len(sys.argv) > 4 else 16
mfd, ne, L = pcb.discretization, pcb.numerical_experiments, pcb._lib
alpha = pcb.dielectric.kpath(d_flag)[0]
relax, pnt = mfd.set_relaxation(alpha)
a_fft, b_fft = mfd.fft_blocks(N, 1, pcb.dielectric.diel_info(d_flag, option="ct"), alpha=alpha)
inv_fft = mfd.inverse_3_times_3_B(b_fft, pnt, relax[0])
Diels = None if typ == "none" else getattr(mfd, typ + "_handle")(N, d_flag)
A, H, P = ne.pc_mfd_handle(a_fft, (pnt * b_fft[0], pnt * b_fft[1]), Diels, inv_fft, relax[0])
ctx = pcb.get_context(N)
X, Y = ctx.random_block(m, 1), ctx.empty(m)
for _ in range(3):
    H.op.apply_into(L.APPLY_H, X, Y)
ctx.sync()
ctx.timer_start()
reps = 20
for _ in range(reps):
    H.op.apply_into(L.APPLY_H, X, Y)
total = ctx.timer_stop() / reps
out = {"N": N, "lattice": d_flag, "type": typ, "cols": m, "ms_per_block_apply": total, "op_applies_per_s": m / total * 1e3,
       "frac_of_336N3_roofline": 336.0 * N ** 3 * m / (total * 1e-3) / 6538.6e9}
buf, npass = (C.c_float * 8)(), C.c_int()
acc = np.zeros(8)
for _ in range(5):
    L.check(L.lib().pcb_apply_timed(H.op.h, L.APPLY_H, m, L.ptr_array(X.ptrs), L.ptr_array(Y.ptrs), buf, C.byref(npass)), "timed")
    acc += np.array(buf[:8])
out["pass_ms"] = [float(v) / 5 for v in acc[:npass.value]]
out["env"] = {k: v for k, v in os.environ.items() if k.startswith("PCB200_") and k != "PCB200_QUIET"}
print(json.dumps(out))
