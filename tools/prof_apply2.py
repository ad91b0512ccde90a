"""ncu driver for the round-2 kernels: python tools/prof_apply2.py <lattice> <type> [N] [cols] -- three H block applies."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PCB200_QUIET", "1")
pcb = importlib.import_module("linear-eigenvalue-problems-in-photonic-crystals_b200")
lattice, typ = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 120
m = int(sys.argv[4]) if len(sys.argv) > 4 else 16
mfd, ne, L = pcb.discretization, pcb.numerical_experiments, pcb._lib
alpha = pcb.dielectric.kpath(lattice)[0]
relax, pnt = mfd.set_relaxation(alpha)
a_fft, b_fft = mfd.fft_blocks(N, 1, pcb.dielectric.diel_info(lattice, option="ct"), alpha=alpha)
inv_fft = mfd.inverse_3_times_3_B(b_fft, pnt, relax[0])
Diels = getattr(mfd, typ + "_handle")(N, lattice)
A, H, P = ne.pc_mfd_handle(a_fft, (pnt * b_fft[0], pnt * b_fft[1]), Diels, inv_fft, relax[0])
ctx = pcb.get_context(N)
X, Y = ctx.random_block(m, 1), ctx.empty(m)
for _ in range(3):
    H.op.apply_into(L.APPLY_H, X, Y)
ctx.sync()
print("done", ctx.launches())
