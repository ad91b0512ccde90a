"""Tiny driver for ncu: a few H block applies (and optionally the LOBPCG block kernels) at the bench shapes.

    python tools/prof_apply.py [N] [cols] [type] [blocks]
"""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PCB200_QUIET", "1")
pcb = importlib.import_module("linear-eigenvalue-problems-in-photonic-crystals_b200")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 120
m = int(sys.argv[2]) if len(sys.argv) > 2 else 16
typ = sys.argv[3] if len(sys.argv) > 3 else "chiral"
blocks = len(sys.argv) > 4 and sys.argv[4] == "blocks"
mfd, ne, L = pcb.discretization, pcb.numerical_experiments, pcb._lib
alpha = pcb.dielectric.kpath("fcc")[0]
relax, pnt = mfd.set_relaxation(alpha)
a_fft, b_fft = mfd.fft_blocks(N, 1, pcb.dielectric.diel_info("fcc", option="ct"), alpha=alpha)
inv_fft = mfd.inverse_3_times_3_B(b_fft, pnt, relax[0])
Diels = getattr(mfd, typ + "_handle")(N, "fcc")
A, H, P = ne.pc_mfd_handle(a_fft, (pnt * b_fft[0], pnt * b_fft[1]), Diels, inv_fft, relax[0])
ctx = pcb.get_context(N)
X, Y = ctx.random_block(m, 1), ctx.empty(m)
for _ in range(2):
    H.op.apply_into(L.APPLY_H, X, Y)
ctx.sync()
if blocks:
    S, HS = ctx.random_block(3 * m, 2), ctx.random_block(3 * m, 3)
    lam = np.linspace(1, 2, m)
    for _ in range(1):
        H.op.residual(S[:, :m], HS[:, :m], S[:, m:2 * m], lam, precond=True)
        pcb.orthogonalization.gram_pair(S, HS)
        E = np.ascontiguousarray(np.random.default_rng(0).standard_normal((3 * m, m)) + 0j) / (3 * m)
        L.check(L.lib().pcb_update(ctx.h, m, 3 * m, L.ptr_array(S.ptrs), L.ptr_array(HS.ptrs), L.ptr_array(S[:, 2 * m:].ptrs),
                                   L.ptr_array(HS[:, 2 * m:].ptrs), E.ctypes.data), "update")
    ctx.sync()
print("done", ctx.launches())
