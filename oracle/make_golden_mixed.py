"""Golden vectors for lobpcg_sep_softlock_mixedprecision (lobpcg.py:494-629) from the UNMODIFIED reference.

    python oracle/make_golden_mixed.py          (build container only: needs /root/reference)

Runs the reference solver through ``oracle/refshim`` on seeded inputs, asserts that the oracle restatement
(``pc_oracle.lobpcg_sep_softlock_mixedprecision``) reproduces it, and writes ``tests/golden/mixedprecision_golden.json``
(eigenvalues, iteration counts, residual histories; inputs are regenerated from the recorded seeds).
"""
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)

import pc_oracle as oc  # noqa: E402
import refload  # noqa: E402

pi = np.pi
OUT = os.path.join(ROOT, "tests", "golden", "mixedprecision_golden.json")
CASES = [("sc_curv", 8, [pi, pi, pi], "chiral", 0, 10, 2000),
         ("bcc_sg", 8, [pi / 2, 0, pi / 3], "pseudochiral_trivial", 0, 10, 2001),
         ("fcc", 12, [pi, 2 * pi, 0], "pseudochiral_crossdof", 0, 6, 2002)]


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    ref = refload.load()
    mfd, lob, ne = ref["discretization"], ref["lobpcg"], ref["numerical_experiments"]
    out = []
    for d_flag, N, alpha, typ, eps_opt, nev, seed in CASES:
        alpha = np.array(alpha, dtype=float)
        a_fft, b_fft, inv_fft, _, shift = quiet(ne.uniform_initialization, N, d_flag, alpha, nev=nev)
        Diels = quiet(getattr(mfd, typ + "_handle"), N, d_flag, eps_opt=eps_opt)
        A_func, H_func, P_func = ne.pc_mfd_handle(a_fft, b_fft, Diels, inv_fft, shift)
        m = nev + round(nev * 0.6)
        x0 = oc.random_x0(3 * N ** 3, m, seed)
        lam, x, info = quiet(lob.lobpcg_sep_softlock_mixedprecision, H_func, P_func, x0.copy(), nev, history=True)
        lam, x, info = np.asarray(lam), np.asarray(x), np.asarray(info)
        assert lam.shape == (nev,) and x.shape == (3 * N ** 3, nev)
        res = np.linalg.norm(np.asarray(H_func(x)) - x * lam, axis=0)
        # oracle restatement on the oracle's own operator
        oa, ob, oi, oshift, _ = oc.assemble_symbols(N, d_flag, alpha)
        diel = getattr(oc, typ + "_handle")(N, d_flag, eps_opt=eps_opt)
        _, oH, oP = oc.pc_mfd_handle(oa, ob, diel, oi, oshift)
        olam, ox, oinfo = oc.lobpcg_sep_softlock_mixedprecision(oH, oP, x0.copy(), nev, history=True)
        assert int(oinfo[0]) == int(info[0]), (d_flag, oinfo[0], info[0])
        e_lam = float(np.max(np.abs(olam - lam) / np.abs(lam)))
        e_his = float(np.max(np.abs(oinfo[2:] - info[2:]) / np.abs(info[2:]))) if len(info) > 2 else 0.0
        assert e_lam < 1e-10 and e_his < 2e-3, (e_lam, e_his)
        out.append({"d_flag": d_flag, "N": N, "alpha": alpha.tolist(), "type": typ, "eps_opt": eps_opt, "nev": nev, "m": m,
                    "seed": seed, "shift": float(shift), "iters": int(info[0]), "lambdas": lam.tolist(),
                    "res_his": info[2:].tolist(), "final_res": res.tolist(),
                    "oracle_vs_reference": {"lambdas_rel": e_lam, "res_his_rel": e_his}})
        print(f"{d_flag} N={N} {typ}: iters={int(info[0])} oracle-vs-reference lam {e_lam:.2e} hist {e_his:.2e}")
    with open(OUT, "w") as f:
        json.dump({"generator": "oracle/make_golden_mixed.py (unmodified reference via oracle/refshim)", "cases": out}, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
