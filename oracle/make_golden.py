"""Generate tests/golden/ from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py

Imports ``/root/reference/paper_2/*.py`` through ``oracle/refshim`` (NumPy-backed
``cupy``/``cupyx``), runs the reference's own functions on seeded inputs and stores
their outputs as small fixtures:

  tests/golden/reference_golden.npz   arrays (symbols, operator outputs, eigenvalues, ...)
  tests/golden/reference_golden.json  manifest: case parameters, seeds, index-set digests

Inputs are never stored: tests regenerate them from the recorded seeds with
``oracle.pc_oracle.random_x0`` (numpy ``default_rng``).  The same script asserts that
the oracle restatement reproduces every stored quantity, i.e. it pins the oracle.
"""
import contextlib
import hashlib
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
OUT_NPZ = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")
OUT_JSON = os.path.join(ROOT, "tests", "golden", "reference_golden.json")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.int64).tobytes()).hexdigest()


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def main():
    ref = refload.load()
    env, diel, mfd, pcfft = ref["environment"], ref["dielectric"], ref["discretization"], ref["pcfft"]
    lob, orth, ne = ref["lobpcg"], ref["orthogonalization"], ref["numerical_experiments"]
    arrays, man = {}, {"index_sets": [], "symbols": [], "operator": [], "rr": [], "lobpcg": [], "stencils": {}}
    worst = {}

    def check(tag, a, b, tol):
        e = relerr(a, b)
        worst[tag] = max(worst.get(tag, 0.0), e)
        assert e <= tol, f"oracle mismatch {tag}: {e:.3e} > {tol}"

    # -- stencils (discretization.py:152-193) --------------------------------------------
    for k in (1, 2, 3):
        for order in (0, 1):
            st = np.asarray(mfd.mfd_stencil(k, order), dtype=float)
            arrays[f"stencil_k{k}_o{order}"] = st
            check("stencil", oc.mfd_stencil(k, order), st, 1e-15)

    # -- index sets (dielectric.py:58-261) -----------------------------------------------
    for d_flag in ("sc_flat1", "sc_flat2", "sc_curv", "bcc_sg", "bcc_dg", "fcc"):
        for N in (6, 8, 12, 16, 24):
            for dofs in ("edge", "volume"):
                ct = diel.diel_info(d_flag, option="ct")
                mesh = quiet(getattr(diel, f"mesh3d_{dofs}_dofs"), N)
                ind = getattr(diel, "FLAG_" + d_flag)(mesh @ np.linalg.inv(ct.T))
                mine = oc.diel_index(N, d_flag, dofs)
                assert np.array_equal(np.asarray(ind), mine), (d_flag, N, dofs)
                man["index_sets"].append({"d_flag": d_flag, "N": N, "dofs": dofs, "count": int(len(ind)),
                                          "sha256": digest(ind)})
    # shipped N>=100 index files: record their digests (regenerable bit-exactly)
    for sub, name in (("volume_dofs", "sc_curv_120"), ("volume_dofs", "fcc_120"), ("volume_dofs", "bcc_dg_120"),
                      ("edge_dofs", "sc_flat1_100")):
        shipped = np.fromfile(os.path.join(refload.REFERENCE, "paper_2", "dielectric_examples", sub, name + ".bin"),
                              dtype=np.int64)
        d_flag, N = name.rsplit("_", 1)
        man["index_sets"].append({"d_flag": d_flag, "N": int(N), "dofs": sub.split("_")[0], "count": int(len(shipped)),
                                  "sha256": digest(shipped), "shipped_file": f"{sub}/{name}.bin"})

    # -- symbols / preconditioner (discretization.py:284-346, numerical_experiments.py:33-71) ----
    sym_cases = [("sc_curv", 6, [pi, pi, pi]), ("sc_curv", 6, [0.0, 0.0, 0.0]), ("fcc", 6, [pi / 20, 0.0, 0.0]),
                 ("bcc_sg", 6, [0.3 * pi, 0.1 * pi, 2 * pi]), ("fcc", 8, [pi, 2 * pi, 0.0])]
    for i, (d_flag, N, alpha) in enumerate(sym_cases):
        alpha = np.array(alpha)
        a_fft, b_fft, inv_fft, x0, shift = quiet(ne.uniform_initialization, N, d_flag, alpha)
        opt, pnt = mfd.set_relaxation(alpha)
        key = f"sym{i}"
        arrays[key + "_a"] = np.asarray(a_fft)
        arrays[key + "_b0"], arrays[key + "_b1"] = np.asarray(b_fft[0]), np.asarray(b_fft[1])
        arrays[key + "_i0"], arrays[key + "_i1"] = np.asarray(inv_fft[0]), np.asarray(inv_fft[1])
        man["symbols"].append({"key": key, "d_flag": d_flag, "N": N, "alpha": alpha.tolist(), "shift": float(shift),
                               "gamma": float(pnt)})
        oa, ob, oi, oshift, ognt = oc.assemble_symbols(N, d_flag, alpha)
        assert oshift == shift and ognt == pnt
        check("a_fft", oa, a_fft, 1e-15)
        check("b_fft0", ob[0], b_fft[0], 1e-15)
        check("b_fft1", ob[1], b_fft[1], 1e-15)
        check("inv_fft0", oi[0], inv_fft[0], 1e-13)
        check("inv_fft1", oi[1], inv_fft[1], 1e-13)
    # bandgap-style assembly (numerical_experiments.py:434-446) must agree with fft_blocks(alpha=...)
    d_fft, di_fft = mfd.fft_blocks(6, env.K, diel.diel_info("fcc", option="ct"))
    od, odi = oc.fft_blocks(6, oc.K, oc.lattice_ct("fcc"))
    check("d_fft", od, d_fft, 1e-15)
    check("di_fft", odi, di_fft, 1e-15)

    # -- operator / preconditioner / dielectric (pcfft.py:130-181, discretization.py:352-453) ----
    op_cases = [("sc_curv", 6, [pi, pi, pi], "chiral", 0, 3), ("fcc", 6, [pi / 20, 0, 0], "chiral", 0, 2),
                ("bcc_sg", 6, [0, 0, 0], "pseudochiral_trivial", 0, 2),
                ("sc_curv", 6, [pi, 0.5 * pi, 0], "pseudochiral_trivial", 3, 2),
                ("bcc_dg", 6, [pi, pi, pi], "pseudochiral_crossdof", 0, 2),
                ("fcc", 6, [pi, 2 * pi, 0], "pseudochiral_crossdof", 3, 2),
                ("sc_curv", 8, [0.2 * pi, 0, 0], "pseudochiral_crossdof", 2, 1),
                ("sc_curv", 6, [pi, pi, pi], None, 0, 2)]
    for i, (d_flag, N, alpha, typ, eps_opt, m) in enumerate(op_cases):
        alpha = np.array(alpha, dtype=float)
        a_fft, b_fft, inv_fft, _, shift = quiet(ne.uniform_initialization, N, d_flag, alpha)
        Diels = (lambda x: x) if typ is None else quiet(getattr(mfd, typ + "_handle"), N, d_flag, eps_opt=eps_opt)
        A_func, H_func, P_func = ne.pc_mfd_handle(a_fft, b_fft, Diels, inv_fft, shift)
        seed = 100 + i
        x = oc.random_x0(3 * N ** 3, m, seed)
        key = f"op{i}"
        hx, ax, px = np.asarray(H_func(x)), np.asarray(A_func(x)), np.asarray(P_func(x))
        dx = np.asarray(Diels(x.copy()))
        h1 = np.asarray(H_func(x[:, 0].copy()))  # 1-D input path (pcfft.py:141-156)
        arrays[key + "_H"], arrays[key + "_A"], arrays[key + "_P"], arrays[key + "_M"] = hx, ax, px, dx
        man["operator"].append({"key": key, "d_flag": d_flag, "N": N, "alpha": alpha.tolist(), "type": typ,
                                "eps_opt": eps_opt, "m": m, "seed": seed, "shift": float(shift)})
        oa, ob, oi, oshift, _ = oc.assemble_symbols(N, d_flag, alpha)
        od = (lambda v: v) if typ is None else oc.HANDLES[typ](N, d_flag, eps_opt=eps_opt)
        oA, oH, oP = oc.pc_mfd_handle(oa, ob, od, oi, oshift)
        check("M", od(x.copy()), dx, 1e-15)
        check("H", oH(x), hx, 1e-14)
        check("A", oA(x), ax, 1e-14)
        check("P", oP(x), px, 1e-13)
        check("H1d", oH(x[:, 0].copy()), h1, 1e-14)

    # -- Rayleigh-Ritz (orthogonalization.py:140-154) ------------------------------------------
    rng = np.random.default_rng(7)
    for i, (R, n) in enumerate([(200, 8), (300, 24)]):
        s = rng.random((R, n)) + 1j * rng.random((R, n))
        q = rng.random((R, R)) + 1j * rng.random((R, R))
        hs = (q + q.conj().T) @ s
        lam, E, _ = orth.rayleigh_ritz_chol_sep(s, hs)
        arrays[f"rr{i}_lam"] = np.asarray(lam)
        arrays[f"rr{i}_absE"] = np.abs(np.asarray(E))
        man["rr"].append({"key": f"rr{i}", "R": R, "n": n, "seed": 7})
        ol, oE = oc.rayleigh_ritz_chol_sep(s, hs)
        check("rr_lam", ol, lam, 1e-12)

    # -- LOBPCG end to end (lobpcg.py:325-492 via numerical_experiments.py:209-247) -------------
    lob_cases = [("sc_curv", 8, [pi, pi, pi], "chiral", 0, 10), ("sc_curv", 12, [pi, pi, pi], "chiral", 0, 10),
                 ("fcc", 8, [pi / 2, 2 * pi, pi / 2], "chiral", 0, 10),
                 ("bcc_sg", 8, [pi / 20, 0, 0], "pseudochiral_trivial", 0, 10),
                 ("sc_curv", 8, [0, 0, 0], "pseudochiral_crossdof", 0, 10),
                 ("bcc_dg", 12, [pi, 0, pi], "pseudochiral_crossdof", 0, 10),
                 ("sc_curv", 16, [pi, pi, pi], "chiral", 0, 10),
                 ("sc_curv", 12, [pi, pi, 0], "chiral", 0, 5)]
    for i, (d_flag, N, alpha, typ, eps_opt, nev) in enumerate(lob_cases):
        alpha = np.array(alpha, dtype=float)
        a_fft, b_fft, inv_fft, _, shift = quiet(ne.uniform_initialization, N, d_flag, alpha, nev=nev)
        Diels = quiet(getattr(mfd, typ + "_handle"), N, d_flag, eps_opt=eps_opt)
        A_func, H_func, P_func = ne.pc_mfd_handle(a_fft, b_fft, Diels, inv_fft, shift)
        m = nev + round(nev * 0.6)
        seed = 1000 + i
        x0 = oc.random_x0(3 * N ** 3, m, seed)
        lam, x, info = quiet(lob.lobpcg_sep_softlock, H_func, P_func, x0.copy(), nev, history=True)
        lam_pnt, lam_re = quiet(ne.recompute_normalize_print, lam[:nev].copy(), x[:, :nev], A_func, shift)
        hx = np.asarray(H_func(x))
        res = np.linalg.norm(hx - np.asarray(x) * lam, axis=0)
        key = f"lob{i}"
        arrays[key + "_lam"] = np.asarray(lam)
        arrays[key + "_info"] = np.asarray(info)
        arrays[key + "_wpnt"], arrays[key + "_wre"] = np.asarray(lam_pnt), np.asarray(lam_re)
        arrays[key + "_res"] = res
        man["lobpcg"].append({"key": key, "d_flag": d_flag, "N": N, "alpha": alpha.tolist(), "type": typ,
                              "eps_opt": eps_opt, "nev": nev, "m": m, "seed": seed, "tol": env.TOL,
                              "iters": int(info[0]), "shift": float(shift)})
        o = oc.eigen_1p(N, d_flag, alpha, type=typ, nev=nev, x0=x0.copy(), eps_opt=eps_opt, history=True)
        assert int(o["info"][0]) == int(info[0]), (key, o["info"][0], info[0])
        check("lob_lam", o["lambdas"][:nev], lam[:nev], 1e-10)
        check("lob_wre", o["omega_re"], lam_re, 1e-10)
        check("lob_hist", o["info"][2:], info[2:], 1e-6)
        print(f"{key}: {d_flag} N={N} {typ} iters={int(info[0])} ok")

    os.makedirs(os.path.dirname(OUT_NPZ), exist_ok=True)
    np.savez_compressed(OUT_NPZ, **arrays)
    man["oracle_vs_reference_worst_relerr"] = worst
    man["generator"] = "oracle/make_golden.py (unmodified reference via oracle/refshim)"
    with open(OUT_JSON, "w") as f:
        json.dump(man, f, indent=1)
    print("oracle vs reference, worst relative errors:")
    for k, v in worst.items():
        print(f"  {k:10s} {v:.3e}")
    print("wrote", OUT_NPZ, os.path.getsize(OUT_NPZ), "bytes")


if __name__ == "__main__":
    main()
