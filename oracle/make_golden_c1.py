"""Golden fixture for BASELINE configs[0] / SURVEY 8(d) "C1" from the UNMODIFIED reference (build container only).

    python oracle/make_golden_c1.py          # ~3 min of CPU

C1 = sc_curv, type "chiral" (eps = 13), N = 48, nev = 10 (m = 16), alpha = (pi, pi, pi), tol 1e-4,
x0 = rng(0).random((R, 16)) + 1j * rng(0).random((R, 16))  (real part drawn first = oracle.random_x0(R, 16, 0)).

Runs ``lobpcg.lobpcg_sep_softlock`` of /root/reference/paper_2 through oracle/refshim (NumPy-backed cupy) and stores
eigenvalues, iteration count, residual history and the post-processed frequencies in tests/golden/c1_golden.json;
asserts that the oracle restatement reproduces them (same iteration count, eigenvalues to 1e-10).
The GPU test (tests/test_gpu_parity_sizes.py::test_c1_solve_vs_reference_golden) solves the same problem from the same x0.
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)

import pc_oracle as oc  # noqa: E402
import refload  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "c1_golden.json")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    N, d_flag, typ, nev, seed = 48, "sc_curv", "chiral", 10, 0
    alpha = np.array([np.pi, np.pi, np.pi])
    ref = refload.load()
    mfd, lob, ne, env = ref["discretization"], ref["lobpcg"], ref["numerical_experiments"], ref["environment"]
    a_fft, b_fft, inv_fft, _, shift = quiet(ne.uniform_initialization, N, d_flag, alpha, nev=nev)
    Diels = quiet(getattr(mfd, typ + "_handle"), N, d_flag)
    A_func, H_func, P_func = ne.pc_mfd_handle(a_fft, b_fft, Diels, inv_fft, shift)
    m = nev + round(nev * 0.6)
    x0 = oc.random_x0(3 * N ** 3, m, seed)
    t0 = time.time()
    lam, x, info = quiet(lob.lobpcg_sep_softlock, H_func, P_func, x0.copy(), nev, history=True)
    t_ref = time.time() - t0
    lam_pnt, lam_re = quiet(ne.recompute_normalize_print, lam[:nev].copy(), x[:, :nev], A_func, shift)
    res = np.linalg.norm(np.asarray(H_func(x)) - np.asarray(x) * lam, axis=0)
    print(f"reference: {int(info[0])} iterations, {t_ref:.1f} s")

    t0 = time.time()
    o = oc.eigen_1p(N, d_flag, alpha, type=typ, nev=nev, x0=x0.copy(), history=True)
    t_or = time.time() - t0
    assert int(o["info"][0]) == int(info[0]), (o["info"][0], info[0])
    err = float(np.max(np.abs(o["lambdas"][:nev] - lam[:nev]) / np.abs(lam[:nev])))
    assert err < 1e-10, err
    print(f"oracle:    {int(o['info'][0])} iterations, {t_or:.1f} s, max rel eigenvalue diff {err:.2e}")

    out = {"generator": "oracle/make_golden_c1.py (unmodified reference via oracle/refshim)",
           "N": N, "d_flag": d_flag, "type": typ, "nev": nev, "m": m, "alpha": alpha.tolist(), "seed": seed,
           "tol": float(env.TOL), "shift": float(shift), "iters": int(info[0]),
           "lambdas": np.asarray(lam).tolist(), "omega_pnt": np.asarray(lam_pnt).tolist(), "omega_re": np.asarray(lam_re).tolist(),
           "residuals": res.tolist(), "res_history": np.asarray(info[2:]).tolist(),
           "reference_cpu_seconds": t_ref, "oracle_cpu_seconds": t_or, "oracle_vs_reference_max_rel_eig": err,
           "cpu_count": os.cpu_count()}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
