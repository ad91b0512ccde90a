"""One-off parity sweep on the GPU: H (and A) against the NumPy oracle for every plane-mode size x dielectric type x stencil
width x eps option (the round-2 pass structures: five-sweep plane pass, coupled M on 3-CTA clusters, cross-DoF plane halves with
the stencil fused / separate, two-tile forward x pass, the z-split plane mode), plus a few five-pass sizes.  Prints one line per case; exits 1 on a miss.

    python tools/sweep_parity.py [max_N]
"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
os.environ.setdefault("PCB200_QUIET", "1")
import pc_oracle as oc  # noqa: E402

pcb = importlib.import_module("linear-eigenvalue-problems-in-photonic-crystals_b200")
mfd, ne = pcb.discretization, pcb.numerical_experiments
oc.FFT_WORKERS = os.cpu_count() or 1
max_n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
TOL = 1e-12


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b))


def case(N, d_flag, typ, k, eps_opt, alpha, options=()):
    ctx = pcb.get_context(N)
    for name, val in options:
        ctx.option(name, val)
    try:
        relax, pnt = mfd.set_relaxation(alpha)
        a_fft, b_fft = mfd.fft_blocks(N, k, pcb.dielectric.diel_info(d_flag, option="ct"), alpha=alpha)
        inv_fft = mfd.inverse_3_times_3_B(b_fft, pnt, relax[0])
        kw = {"eps_opt": eps_opt}
        if typ == "pseudochiral_crossdof":
            kw["k"] = k
        Diels = None if typ is None else getattr(mfd, typ + "_handle")(N, d_flag, **kw)
        A, H, P = ne.pc_mfd_handle(a_fft, (pnt * b_fft[0], pnt * b_fft[1]), Diels, inv_fft, relax[0])
        ao, bo, io, shift, _ = oc.assemble_symbols(N, d_flag, alpha, k=k)
        okw = {"ind_e": pcb.dielectric.compute_index(N, d_flag, "edge")}
        if typ == "pseudochiral_trivial":
            okw["ind_v"] = pcb.dielectric.compute_index(N, d_flag, "volume")
        if typ == "pseudochiral_crossdof":
            okw["k"] = k
        diel = (lambda v: v) if typ is None else oc.HANDLES[typ](N, d_flag, eps_opt=eps_opt, **okw)
        Ao, Ho, Po = oc.pc_mfd_handle(ao, bo, diel, io, shift)
        x = oc.random_x0(3 * N ** 3, 2, N + k)
        eh, ea = relerr(H(x), Ho(x)), relerr(A(x), Ao(x))
    finally:
        for name, val in options:
            ctx.option(name, {"mid_five": -1, "plane_cross": 1, "plane_coupled": 1, "plane": 1, "plane_split": 0}[name])
    return eh, ea


def main():
    bad = 0
    t0 = time.time()
    sizes = [n for n in (8, 16, 24, 32, 48, 64, 72, 80, 96, 120, 12, 100, 128, 144, 160) if n <= max_n]
    split_sizes = (16, 32, 48, 64, 96)      # both forms of the plane mode exist: also run the z-split form (the default of 128, 144, 160)
    lattices = {"chiral": "fcc", None: "sc_curv", "pseudochiral_trivial": "bcc_sg", "pseudochiral_crossdof": "bcc_dg"}
    alphas = [np.array([np.pi, 0.3 * np.pi, 0.0]), np.array([0.0, 0.0, 2 * np.pi])]
    n_cases = 0
    for N in sizes:
        for typ in ("chiral", None, "pseudochiral_trivial", "pseudochiral_crossdof"):
            eps_opts = (0, 3) if typ in ("pseudochiral_trivial", "pseudochiral_crossdof") else (0,)
            ks = (1, 2) if (typ == "pseudochiral_crossdof" and N <= 48) else (1,)
            for eps_opt in eps_opts:
                for k in ks:
                    variants = [()]
                    if typ == "pseudochiral_crossdof" and N % 8 == 0:
                        variants.append((("plane_cross", 2),))
                    if typ in ("chiral", None, "pseudochiral_trivial") and N in (24, 72):
                        variants.append((("mid_five", 1),))
                    if N in split_sizes:
                        variants.append((("plane_split", 1),))
                        if typ == "pseudochiral_crossdof":
                            variants.append((("plane_split", 1), ("plane_cross", 2)))
                    if N > 120 and typ == "pseudochiral_crossdof":
                        variants = [(("plane_cross", 1),), (("plane_cross", 2),)]
                    for opts in variants:
                        alpha = alphas[n_cases % 2]
                        eh, ea = case(N, lattices[typ], typ, k, eps_opt, alpha, opts)
                        ok = eh < TOL and ea < TOL
                        bad += 0 if ok else 1
                        n_cases += 1
                        print(f"N={N:3d} {str(typ):22s} eps_opt={eps_opt} k={k} {dict(opts)!s:40s} H {eh:.2e} A {ea:.2e} {'ok' if ok else 'MISS'}", flush=True)
    print(f"{n_cases} cases, {bad} misses, {time.time() - t0:.0f} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
