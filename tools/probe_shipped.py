"""Which discretisation produced the reference's shipped N=120 band rows?  Try stencil half-width k = 1, 2, 3."""
import json, os, sys, importlib, numpy as np
os.environ["PCB200_QUIET"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pcb = importlib.import_module("linear-eigenvalue-problems-in-photonic-crystals_b200")
mfd, ne = pcb.discretization, pcb.numerical_experiments
rows = json.load(open("tests/golden/shipped_bands.json"))["rows"]
np.set_printoptions(precision=8, linewidth=200)
for row in rows:
    if (row["type"], row["d_flag"], row["k_index"]) not in (("chiral", "fcc", 0), ("chiral", "sc_curv", 59), ("pseudochiral_trivial", "sc_curv", 59)):
        continue
    N, d_flag = 120, row["d_flag"]
    alpha = pcb.dielectric.kpath(d_flag)[row["k_index"]]
    want = np.array(row["frequencies"][:10])
    print(row["type"], d_flag, row["k_index"], "alpha/pi", alpha / np.pi)
    print("  shipped", want)
    for k in (1, 2):
        relax, pnt = mfd.set_relaxation(alpha)
        a_fft, b_fft = mfd.fft_blocks(N, k, pcb.dielectric.diel_info(d_flag, option="ct"), alpha=alpha)
        inv_fft = mfd.inverse_3_times_3_B(b_fft, pnt, relax[0])
        Diels = getattr(mfd, row["type"] + "_handle")(N, d_flag)
        A, H, P = ne.pc_mfd_handle(a_fft, (pnt * b_fft[0], pnt * b_fft[1]), Diels, inv_fft, relax[0])
        x0 = pcb.get_context(N).random_block(16, 5)
        lam, x, info = pcb.lobpcg.lobpcg_sep_softlock(H, P, x0, 10)
        w_pnt, w_re = ne.recompute_normalize_print(lam[:10], x[:, :10], A, relax[0])
        print(f"  k={k} iters {int(info[0])}", w_re, "maxdiff %.2e" % np.max(np.abs(w_re - want)), flush=True)
