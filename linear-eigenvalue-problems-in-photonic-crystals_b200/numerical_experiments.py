"""Runners of the hot path (public surface of paper_2/numerical_experiments.py).

uniform_initialization :33-71, pc_mfd_handle :73-85, recompute_normalize_print :87-158, eigen_1p :209-247
and bandgap :313-496 keep their names, arguments, printed tables and the JSON checkpoint format
("<d_flag>_<n>_iterations" / "_frequencies", [0,0] = uncomputed, [-1,-1] = failed).  New here:
``kpath_chunks`` / ``bandgap_sharded`` distribute the k-path over the GPUs of a node, one k-chunk per
process, with no data-path collective.
"""
import json
import os
import time

import numpy as np
from numpy import pi

from . import _lib as L
from . import devarray
from . import dielectric as diel
from . import discretization as mfd
from .devarray import DeviceBlock
from .environment import (CYAN, GAP, GREEN, K, NEV, OUTPUT_PATH, RED, RESET, SCAL, TOL, TYPE0, YELLOW, owari_cuda, say)
from .lobpcg import lobpcg_sep_softlock, _residual_helper
from .pcfft import OperatorCallable, build_operator, column_dots

owari = owari_cuda
_seed_counter = [None]


def _next_seed(seed=None):
    if seed is not None:
        return int(seed)
    if _seed_counter[0] is None:
        _seed_counter[0] = int(np.random.SeedSequence().entropy % (1 << 62))
    _seed_counter[0] += 1
    return _seed_counter[0]


def _handle(type, n, d_flag, **kw):
    if type is None:
        return None
    fn = getattr(mfd, type + "_handle", None)
    if fn is None:
        raise ValueError(f"unknown dielectric type {type!r}")
    return fn(n, d_flag, **kw)


def uniform_initialization(n, d_flag, alpha, nev=NEV, k=K, seed=None, device=None):
    """Symbols, initial guess and shift for one lattice vector (numerical_experiments.py:33-71).
    Returns (a_fft, b_fft, inv_fft, x0, shift); the symbols are O(N) descriptors, x0 a device block."""
    t_h = time.time()
    alpha = np.asarray(alpha, dtype=float)
    relax_opt, pnt = mfd.set_relaxation(alpha, scal=SCAL)
    ct = diel.diel_info(d_flag, option="ct")
    a_fft, b_fft = mfd.fft_blocks(n, k, ct, alpha=alpha, scal=SCAL)
    inv_fft = mfd.inverse_3_times_3_B(b_fft, pnt, relax_opt[0])
    a_fft = a_fft / SCAL
    b_fft = (pnt * b_fft[0] / SCAL / SCAL, pnt * b_fft[1] / SCAL / SCAL)
    inv_fft = (inv_fft[0] * SCAL * SCAL, inv_fft[1] * SCAL * SCAL)
    m = round(nev * relax_opt[1]) + nev
    x0 = devarray.get_context(n, device).random_block(m, _next_seed(seed))
    say(f"Matrix blocks done, {owari() - t_h:<6.3f}s elapsed.")
    return a_fft, b_fft, inv_fft, x0, relax_opt[0]


def pc_mfd_handle(a_fft, b_fft, Diels, inv_fft, shift=0.0):
    """(A_func, H_func, P_func): AMA', AMA' + gamma B'B + shift, inv(AA' + gamma B'B + shift)
    (numerical_experiments.py:73-85) as callables over one device operator."""
    op = build_operator(a_fft, b_fft, Diels, inv_fft, shift)
    return OperatorCallable(op, L.APPLY_A), OperatorCallable(op, L.APPLY_H), OperatorCallable(op, L.APPLY_P)


def _sqrt_robust(a):
    return 0.0 if (a <= 0) and (a > -1e-8) else a ** 0.5


def recompute_normalize_print(lambdas_in, x, A_func, shift=0.0, scal=SCAL):
    """Recompute eigenvalues without penalty/shift, print the table, flag spurious modes
    (numerical_experiments.py:87-158).  Returns (omega_pnt, omega_re) as frequencies omega/2pi."""
    t_h = time.time()
    lambdas_in = np.asarray(lambdas_in, dtype=float)
    if not isinstance(x, DeviceBlock):
        n = round((np.asarray(x).shape[0] // 3) ** (1 / 3))
        x = DeviceBlock.from_host(devarray.get_context(n), x)
    if isinstance(A_func, OperatorCallable) and A_func.op.ctx is x.ctx:
        adax = A_func.op.apply_into(A_func.mode, x, x.ctx.work_block("recompute.adax", x.k))   # cached work space
    else:
        adax = A_func(x)
    lambdas_pnt = lambdas_in - shift if shift > 0.0 else lambdas_in.copy()
    tmp = x.ctx.work_block("recompute.tmp", x.k)
    res = _residual_helper(x.ctx).residual(x, adax, tmp, lambdas_pnt, precond=False)    # ||x lambda - A x||
    lambdas_re = (column_dots(x, adax) / column_dots(x, x)).real
    for i in np.where(np.isnan(lambdas_pnt))[0]:
        if np.isnan(lambdas_re[i]):
            say(f"{RED}Warning: NaN occurs in both lambda_pnt and lambda_re, index = {i}. Please run the program again.")
        else:
            say(f"{YELLOW}Warning: NaN occurs in lambda_pnt, index = {i}, but same index in lambda_re is valid.")
    for i in np.where(np.isnan(lambdas_re))[0]:
        if not np.isnan(lambdas_pnt[i]):
            lambdas_re[i] = lambdas_pnt[i]
    say(f"Runtime for recomputing: {owari() - t_h:<6.3f}s.")
    say("| i  |    omega   |  omega_re  |  abs(omega - omega_re)  | residual  |")
    flag_spurious = False
    for i in range(len(lambdas_pnt)):
        l1 = _sqrt_robust(lambdas_pnt[i]) * scal / (2 * pi)
        l2 = _sqrt_robust(lambdas_re[i]) * scal / (2 * pi)
        say(f"| {i + 1:<2d} | {l1:<10.6f} | {l2:<10.6f} |        {abs(l1 - l2):<10.3e}       | {res[i]:<6.3e} |")
        lambdas_pnt[i], lambdas_re[i] = l1, l2
        if l1 - l2 > 1e-3:
            flag_spurious = True
    if flag_spurious:
        raise ValueError(f"{RED}Spurious eigenvalues occur.{RESET}")
    recompute_normalize_print.last_residuals = res
    return lambdas_pnt, lambdas_re


def eigen_1p(n, d_flag, alpha, type="chiral", nev=NEV, solver=lobpcg_sep_softlock, cpu_x0=False, x0=None, seed=None,
             tol=TOL, eps_opt=0):
    """Eigenvalues at one lattice vector (numerical_experiments.py:209-247).  Prints like the reference and
    additionally returns a dict (lambdas, x, info, omega_pnt, omega_re, residuals, shift)."""
    alpha = np.asarray(alpha, dtype=float)
    a_fft, b_fft, inv_fft, x_rand, shift = uniform_initialization(n, d_flag, alpha, nev=nev, seed=seed)
    if x0 is None:
        x0 = x_rand
    del x_rand
    if cpu_x0:
        (x0.get() if isinstance(x0, DeviceBlock) else np.asarray(x0)).tofile("x0_from_single_gpu.bin")
    Diels = _handle(type, n, d_flag, eps_opt=eps_opt) if type is not None else None
    A_func, H_func, P_func = pc_mfd_handle(a_fft, b_fft, Diels, inv_fft, shift)
    lambdas_pnt, x, iters = solver(H_func, P_func, x0, nev, tol=tol)
    del x0
    if lambdas_pnt is None:
        say(f"{RED}Solver failed (NaN / blow-up).{RESET}")
        return None
    say(f"n = {n}, lattice type: {d_flag}, alpha = [{alpha[0] / pi:<5.2f}, {alpha[1] / pi:<5.2f}, {alpha[2] / pi:<5.2f}] pi, "
        f"iter = {int(iters[0])}, runtime = {iters[1]:<6.3f}s.")
    say(f"{CYAN}\nEigenvalues (not sqrt normalized) are:\n")
    say(lambdas_pnt)
    say(f"\n{RESET}")
    w_pnt, w_re = recompute_normalize_print(lambdas_pnt[:nev], x[:, :nev], A_func, shift=shift)
    return {"lambdas": lambdas_pnt, "x": x, "info": iters, "omega_pnt": w_pnt, "omega_re": w_re,
            "residuals": recompute_normalize_print.last_residuals, "shift": shift}


# ---------------------------------------------------------------------------------------------
# Band structure along the k-path
# ---------------------------------------------------------------------------------------------
def _load_or_init_record(path_bandgap, var_it, var_fq, n_k, type, d_flag, n):
    """JSON checkpoint logic of bandgap() (numerical_experiments.py:355-408).
    Returns (gap_lib, uncomputed indices or None, done flag)."""
    fresh_it, fresh_fq = [[0] * 2 for _ in range(n_k)], [[0] * NEV for _ in range(n_k)]
    if not os.path.exists(path_bandgap):
        say("The bandgap of type ", d_flag, " has no previous record.")
        gap_lib = {var_it: fresh_it, var_fq: fresh_fq}
        os.makedirs(os.path.dirname(path_bandgap) or ".", exist_ok=True)
        with open(path_bandgap, "w") as f:
            json.dump(gap_lib, f, indent=4)
        return gap_lib, None, False
    with open(path_bandgap, "r") as f:
        gap_lib = json.load(f)
    if var_it in gap_lib:
        say(f"{GREEN}Lattice type {type},{d_flag} with grid size n = {n} has a_fft previous record.{RESET}")
        rec = gap_lib[var_it]
        err_ind = [i for i, a in enumerate(rec) if a == [-1, -1]]
        if err_ind:
            say(f"{RED}Warning: Blow up results detected: {err_ind}.{RESET}")
        empty_ind = [i for i, a in enumerate(rec) if a == [0, 0]]
        if empty_ind:
            say(f"{YELLOW}Following indices remain uncomputed: {empty_ind}.{RESET}")
        if not empty_ind and not err_ind:
            say(f"{GREEN}All indices of {type},{d_flag} have been computed without errors.{RESET}")
            return gap_lib, [], True
        return gap_lib, sorted(set(err_ind + empty_ind)), False
    say(f"{YELLOW}Lattice type {type},{d_flag} will be computed with a_fft new grid size n = {n}.{RESET}")
    gap_lib[var_it], gap_lib[var_fq] = fresh_it, fresh_fq
    with open(path_bandgap, "w") as f:
        json.dump(gap_lib, f, indent=4)
    return gap_lib, None, False


def bandgap(n, d_flag, solver=lobpcg_sep_softlock, type=TYPE0, eps_opt=0, indices=None, nev=NEV, seed=None,
            path=None, tol=TOL / SCAL / SCAL, only=None):
    """Band structure along the lattice's k-path, one warm-started LOBPCG solve per point, checkpointed to
    JSON after every point (numerical_experiments.py:313-496).  Returns the list of failed indices.
    `nev` (default NEV), `seed` (reproducible random starts) and `path` (output file) are additions."""
    t_all = time.time()
    timing = {"setup": 0.0, "assemble": 0.0, "solver": 0.0, "postprocess": 0.0, "checkpoint": 0.0}
    bandgap.last_timing = timing          # wall seconds by phase (where a band-structure run spends its time)
    ct, sym_points = diel.diel_info(d_flag)
    alphas = diel.kpath(d_flag, GAP)
    n_k = alphas.shape[0]
    Diels = _handle(type, n, d_flag, eps_opt=eps_opt)
    ctx = Diels.ctx if Diels is not None else devarray.get_context(n)
    d_fft, _ = mfd.fft_blocks(n, K, ct)
    timing["setup"] = time.time() - t_all

    path_bandgap = path or (OUTPUT_PATH + type + "/bandgap_" + d_flag + ".json")
    var_it, var_fq = f"{d_flag}_{n}_iterations", f"{d_flag}_{n}_frequencies"
    gap_lib, uncomputed, done = _load_or_init_record(path_bandgap, var_it, var_fq, n_k, type, d_flag, n)
    if done:
        return []
    gap_rec_it, gap_rec_fq = gap_lib[var_it], gap_lib[var_fq]
    if indices is None:
        indices = list(range(n_k)) if uncomputed is None else uncomputed
        if only is not None:          # resume restricted to the given rows (the file may hold other ranks' empty rows)
            indices = [i for i in indices if i in set(only)]
            if not indices:
                return []
    elif len(indices) == 0:
        return []
    indices = [int(i) for i in indices]
    if max(indices) >= n_k or min(indices) < 0:
        raise ValueError("Index is non-positive or is incompatible with gap size.")

    err_index, op, x = [], None, None
    for i, idx in enumerate(indices):
        t_h = time.time()
        alpha = alphas[idx] / SCAL
        relax_opt, pnt = mfd.set_relaxation(alpha)
        m = nev + round(nev * relax_opt[1])
        if i == 0 or abs(indices[i] - indices[i - 1]) > 1 or x is None:
            x0 = ctx.random_block(m, _next_seed(None if seed is None else seed + idx))
        elif m <= x.shape[1]:
            x0 = x[:, 0:m]
        else:
            x0 = ctx.empty(m)
            x0[:, :x.shape[1]] = x
            x0[:, x.shape[1]:] = ctx.random_block(m - x.shape[1], _next_seed(None if seed is None else seed + idx))
        a_fft = d_fft.with_alpha(alpha)
        b_fft = mfd.PenaltySymbols(a_fft)
        inv_fft = mfd.inverse_3_times_3_B(b_fft, pnt, relax_opt[0])
        a_fft = a_fft / SCAL
        b_fft = (pnt * b_fft[0] / SCAL / SCAL, pnt * b_fft[1] / SCAL / SCAL)
        inv_fft = (inv_fft[0] * SCAL * SCAL, inv_fft[1] * SCAL * SCAL)
        A_func, H_func, P_func = pc_mfd_handle(a_fft, b_fft, Diels, inv_fft, relax_opt[0])
        say(f"Matrix blocks done, {owari() - t_h:<6.3f}s elapsed.")
        timing["assemble"] += time.time() - t_h
        try:
            t_s = time.time()
            lambdas_pnt, x, iters = solver(H_func, P_func, x0, nev, tol=tol)
            timing["solver"] += time.time() - t_s
            if lambdas_pnt is None:
                raise RuntimeError("solver returned None (NaN / blow-up)")
            say(f"Gap {idx + 1} out of {n_k} ({d_flag}),"
                f"alpha = ({alpha[0] / pi:<6.3f}, {alpha[1] / pi:<6.3f}, {alpha[2] / pi:<6.3f})pi is computed.")
            say(f"Iterations = {int(iters[0])}, runtime = {iters[1]:<6.3f}s.\n")
            t_s = time.time()
            _, lambdas_re = recompute_normalize_print(lambdas_pnt, x, A_func, relax_opt[0])
            timing["postprocess"] += time.time() - t_s
            gap_rec_it[idx] = [float(v) for v in iters[:2]]
            gap_rec_fq[idx] = [float(v) for v in lambdas_re]
        except Exception as e:          # same policy as the reference: record [-1,-1], restart from random
            say(f"{RED}WARNING: Error occurs.")
            say(f"Error message: {e}{RESET}")
            err_index.append(idx)
            x = None
            gap_rec_it[idx] = [-1.0, -1.0]
            gap_rec_fq[idx] = [-1.0] * nev
        del x0
        t_h = time.time()
        gap_lib[var_it], gap_lib[var_fq] = gap_rec_it, gap_rec_fq
        with open(path_bandgap, "w") as f:
            json.dump(gap_lib, f, indent=4)
        timing["checkpoint"] += time.time() - t_h
        say(f"{CYAN}Gap info library ({d_flag}) is updated ({idx + 1}/{n_k}), time = {time.time() - t_h:<6.3f}s.{RESET}")
    if err_index:
        say(f"{RED}Error occurs to following indices:{RESET}")
        say(err_index)
    else:
        say(f"{GREEN}All indices computed correctly.{RESET}")
    timing["total"] = time.time() - t_all
    return err_index


def kpath_chunks(n_k, world):
    """Contiguous k-chunks, one per GPU, so the warm-start chain x0 <- x survives inside a chunk."""
    base, extra = divmod(n_k, world)
    out, start = [], 0
    for r in range(world):
        size = base + (1 if r < extra else 0)
        out.append(list(range(start, start + size)))
        start += size
    return out


def merge_records(parts, n_k, nev=NEV):
    """Combine per-rank {"iterations": {idx: [it, s]}, "frequencies": {idx: [...]}} rows into full-path lists."""
    its, fqs = [[0, 0] for _ in range(n_k)], [[0] * nev for _ in range(n_k)]
    for p in parts:
        for idx, row in p["iterations"].items():
            its[int(idx)] = row
        for idx, row in p["frequencies"].items():
            fqs[int(idx)] = row
    return its, fqs


def bandgap_sharded(n, d_flag, rank, world, type=TYPE0, eps_opt=0, nev=NEV, seed=1000, out_dir=None, indices=None,
                    gather=None, tol=TOL / SCAL / SCAL, retry=1):
    """k-path sharding over `world` processes (one GPU each): rank r solves chunk r of the path (or of `indices`)
    into its own JSON file; with `gather` (callable: obj -> list of objs on rank 0, e.g. a torch.distributed
    gather_object wrapper) rank 0 merges all rows into the reference-format file.  No collective touches the data path."""
    all_idx = list(range(diel.kpath(d_flag, GAP).shape[0])) if indices is None else list(indices)
    mine = [all_idx[i] for i in kpath_chunks(len(all_idx), world)[rank]]
    out_dir = out_dir or (OUTPUT_PATH + type + "/")
    os.makedirs(out_dir, exist_ok=True)
    part_path = os.path.join(out_dir, f"bandgap_{d_flag}.rank{rank}.json")
    if os.path.exists(part_path):
        os.remove(part_path)
    errs = bandgap(n, d_flag, type=type, eps_opt=eps_opt, indices=mine, nev=nev, seed=seed, path=part_path, tol=tol) if mine else []
    # A warm start can make [X W] numerically rank deficient (e.g. the point after Gamma, whose zero modes are pure gradients):
    # the reference records [-1,-1] and recomputes such rows from a random start when bandgap() is called again
    # (numerical_experiments.py:370-390); do that second call here.
    for attempt in range(retry):
        if not errs:
            break
        errs = bandgap(n, d_flag, type=type, eps_opt=eps_opt, indices=None, nev=nev, seed=seed + 7919 * (attempt + 1),
                       path=part_path, tol=tol, only=errs)
    rows = {"iterations": {}, "frequencies": {}, "errors": errs}
    if mine:
        with open(part_path) as f:
            lib = json.load(f)
        for idx in mine:
            rows["iterations"][idx] = lib[f"{d_flag}_{n}_iterations"][idx]
            rows["frequencies"][idx] = lib[f"{d_flag}_{n}_frequencies"][idx]
    if gather is None:
        return rows
    parts = gather(rows)
    if rank == 0:
        n_k = diel.kpath(d_flag, GAP).shape[0]
        its, fqs = merge_records(parts, n_k, nev)
        final = os.path.join(out_dir, f"bandgap_{d_flag}.json")
        lib = {}
        if os.path.exists(final):
            with open(final) as f:
                lib = json.load(f)
        lib[f"{d_flag}_{n}_iterations"], lib[f"{d_flag}_{n}_frequencies"] = its, fqs
        with open(final, "w") as f:
            json.dump(lib, f, indent=4)
    return rows
