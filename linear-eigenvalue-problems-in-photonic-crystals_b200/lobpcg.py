"""Soft-locking LOBPCG on the GPU (public surface of paper_2/lobpcg.py for the default solver).

``lobpcg_sep_softlock`` keeps the reference's signature, return convention and failure behaviour
(lobpcg.py:325-492) and reproduces its iteration exactly -- including the initial-lambda quirk
(:379-381), the iteration-0 search space without P, the absolute residual test and the ascending
soft-lock order -- but every O(R) step is one fused kernel of libpcb200.so:

    residual + column norms + preconditioner      pcb_residual   (lobpcg.py:394-397,442; first iteration only -- afterwards
                                                   fused into the update of the previous iteration, pcb_update_resid)
    soft-lock "compaction"                         pointer lists  (lobpcg.py:431-436: no data moves)
    H on the active block                          pcb_apply      (lobpcg.py:443)
    Gram pair S^H S, S^H HS                        pcb_gram2      (orthogonalization.py:143-144)
    P <- [W P] E, X <- X E + P (and HS)            pcb_update     (lobpcg.py:1248-1270)

Host Python owns the control flow and the n_loc x n_loc Rayleigh-Ritz (LAPACK via NumPy).
"""
import os
import time

import numpy as np

from . import _lib as L
from . import devarray
from .devarray import DeviceBlock
from .environment import GREEN, MAXITER, RED, RESET, TOL, YELLOW, say
from .orthogonalization import gram_pair, gram_pair_top, hermitize, rr_small
from .pcfft import Operator, OperatorCallable


def _small_blas_threads():
    """Context manager: one BLAS thread while the solver runs.  Its host linear algebra is n_loc x n_loc (<= 96): OpenBLAS'
    thread hand-off costs more than the products themselves (48 x 48: two matmuls 0.20 ms on 8 threads, 0.06 ms on one), and the
    GPU idles meanwhile.  No-op without threadpoolctl."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=1, user_api="blas")
    except Exception:
        import contextlib
        return contextlib.nullcontext()


def _fused_operator(h_func_in, p_func, shift):
    """The Operator when (h_func, p_func) are the callables of pc_mfd_handle for one operator."""
    if (shift == 0.0 and isinstance(h_func_in, OperatorCallable) and isinstance(p_func, OperatorCallable)
            and h_func_in.op is p_func.op and h_func_in.mode == L.APPLY_H and p_func.mode == L.APPLY_P):
        return h_func_in.op
    return None


def _call_into(func, src, dst):
    """dst <- func(src) for a generic callable over DeviceBlocks."""
    out = func(src)
    if not isinstance(out, DeviceBlock):
        out = DeviceBlock.from_host(src.ctx, out)
    dst.assign(out)


def _context_of(h_func_in, x0):
    if isinstance(x0, DeviceBlock):
        return x0.ctx
    if isinstance(h_func_in, OperatorCallable):
        return h_func_in.op.ctx
    n = round((x0.shape[0] // 3) ** (1 / 3))
    return devarray.get_context(n)


# Incremental Gram pair.  Between two iterations only the block W is new: X' = S E and P' = S_WP E_WP are rotations of the
# previous search space, so their Gram blocks follow from the previous small matrices ([E E_p]^H G [E E_p], same for T) and
# only the rows of W need an O(R) pass (pcb_gram2_top): 45 % fewer flops at n_act = m, more as columns lock.  The blocks are
# mathematically those the reference forms with two full ZGEMMs (orthogonalization.py:143-144); every GRAM_REFRESH-th
# iteration the full pair is recomputed from the vectors, which bounds the accumulated rounding drift.  PCB200_FULL_GRAM=1
# (or incremental_gram=False) recomputes it every iteration.
GRAM_REFRESH = 8


def lobpcg_sep_softlock(h_func_in, p_func, x0, nev, shift=0.0, tol=TOL, maxiter=MAXITER, history=False,
                        longortho=False, singleprecision=False, maxstagniter=50, trace=None, _lock=True,
                        incremental_gram=None, _mixed=False):
    with _small_blas_threads():
        return _lobpcg_sep_softlock(h_func_in, p_func, x0, nev, shift, tol, maxiter, history, longortho, singleprecision,
                                    maxstagniter, trace, _lock, incremental_gram, _mixed)


lobpcg_sep_softlock.__doc__ = """LOBPCG with soft locking (lobpcg.py:325-492); see _lobpcg_sep_softlock."""


def _lobpcg_sep_softlock(h_func_in, p_func, x0, nev, shift=0.0, tol=TOL, maxiter=MAXITER, history=False,
                         longortho=False, singleprecision=False, maxstagniter=50, trace=None, _lock=True,
                         incremental_gram=None, _mixed=False):
    """LOBPCG with soft locking; [X, W, P] and their images live in two 3m-column device blocks.

    Returns ``(lambdas[:m] - shift, x, info)`` with ``x`` a DeviceBlock (R x m), ``info = [iterations,
    seconds]`` (+ residual history when ``history``); ``(None, None, None)`` on NaN / Cholesky failure /
    blow-up, exactly like the reference.  ``trace`` (list) receives per-iteration dicts for tests."""
    if longortho or singleprecision:
        raise NotImplementedError("longortho / singleprecision variants are outside the ported hot path "
                                  "(SURVEY.md 2.2: not used by any default runner)")
    t_h = time.time()
    ctx = _context_of(h_func_in, x0)
    m = x0.shape[1]
    op = _fused_operator(h_func_in, p_func, shift)
    if shift == 0.0:
        h_func = h_func_in
    else:
        def h_func(x):
            y = h_func_in(x)
            L.check(L.lib().pcb_axpby(ctx.h, x.k, L.ptr_array(x.ptrs), L.ptr_array(y.ptrs), float(shift), 1.0), "pcb_axpby")
            return y
    res_op = op if op is not None else _residual_helper(ctx)

    S, HS = ctx.work_block("lobpcg.S", 3 * m), ctx.work_block("lobpcg.HS", 3 * m)     # cached across solves
    X, W, P = S[:, :m], S[:, m:2 * m], S[:, 2 * m:]
    HX, HW, HP = HS[:, :m], HS[:, m:2 * m], HS[:, 2 * m:]
    if isinstance(x0, DeviceBlock):
        X.assign(x0)
    else:
        X.set(np.asarray(x0))

    if op is not None:
        op.apply_into(L.APPLY_H, X, HX)
    else:
        _call_into(h_func, X, HX)
    ss, shs = gram_pair(X, HX)
    if _mixed:
        # mixed-precision variant: plain eigenvalues of herm(X^H HX) (lobpcg.py:524)
        lambdas = np.linalg.eigvalsh(shs)
    else:
        # Initial lambda (lobpcg.py:379-381): the m x m Gram matrices are themselves fed to RR.
        try:
            lambdas, _ = rr_small(hermitize(ss.conj().T @ ss), hermitize(ss.conj().T @ shs))
        except np.linalg.LinAlgError:
            return None, None, None
    res_his = np.empty(maxiter)
    ctx.sync()
    say(f"Time for LOBPCG initialization: {time.time() - t_h:<6.2f}s.")

    if incremental_gram is None:
        incremental_gram = os.environ.get("PCB200_FULL_GRAM", "0") != "1"
    g_xp = t_xp = None        # Gram pair of [X | P] (2m x 2m) implied by the last Rayleigh-Ritz rotation
    # The residual of iteration i+1 only needs the Ritz values and the X, HX that the update of iteration i produces, so that
    # update also forms it (from its accumulators: X and HX are not read again), applies K_P^-1 and returns the norms.
    fuse_resid = op is not None and not _mixed and os.environ.get("PCB200_FUSED_RESID", "1") != "0"
    next_nrms = None
    t_tot_h = time.time()
    iter_ = 0
    for iter_ in range(maxiter):
        t_iter_h = time.time()
        # residual (+ preconditioner on the fused path), norms, active set
        if next_nrms is not None:
            res_nrms, next_nrms = next_nrms, None
        else:
            res_nrms = res_op.residual(X, HX, W, lambdas[:m], precond=op is not None, single=_mixed)
        res_his[iter_] = np.linalg.norm(res_nrms[:nev])
        ind_act = np.where(res_nrms > tol)[0] if _lock else np.arange(m)
        n_act = len(ind_act)
        if trace is not None:
            trace.append({"res": res_nrms.copy(), "n_act": n_act, "lambdas": np.array(lambdas[:m])})
        say(f"Iter = {iter_:<4d}, res_nrm = {np.linalg.norm(res_nrms):<6.2e}, n_act = {n_act:<3d}.", end=" ")
        if np.isnan(res_nrms).any():
            say(f"{RED}Nan occurs in residuals.{RESET}")
            if not _lock or _mixed:
                raise ValueError(f"{RED}Nan occurs in residuals.{RESET}")      # lobpcg_sep_nolock raises (lobpcg.py:139-140,549)
            return None, None, None
        if (iter_ > maxstagniter and (res_nrms[0] > 1000 or res_nrms[0] > res_his[1])) or \
                (iter_ > 2 * maxstagniter and res_nrms[0] > 50):
            if np.linalg.norm(res_nrms[:nev]) < res_his[maxstagniter // 2] * 0.1:
                say(f"{YELLOW}Stagnation warning.{RESET}")
            else:
                say(f"{YELLOW}Stagnation detected, probably blowup but no nan occurs.{RESET}")
                return None, None, None
        if max(res_nrms[:nev]) < tol:
            say(f"{GREEN}convergence reached.{RESET}")
            break
        n_loc = m + 2 * n_act if iter_ > 0 else m + n_act

        # soft locking = column views of the active W / P / HP (no copies)
        W_act, HW_act = W.cols(ind_act), HW.cols(ind_act)
        if op is not None:
            op.apply_into(L.APPLY_H, W_act, HW_act)          # W already holds K_P^-1 r
        else:
            _call_into(p_func, W_act, W_act)
            _call_into(h_func, W_act, HW_act)
        if iter_ > 0:
            s_loc = DeviceBlock(ctx, _owners=S._owners, _ptrs=X.ptrs + W_act.ptrs + P.cols(ind_act).ptrs)
            hs_loc = DeviceBlock(ctx, _owners=HS._owners, _ptrs=HX.ptrs + HW_act.ptrs + HP.cols(ind_act).ptrs)
        else:
            s_loc = DeviceBlock(ctx, _owners=S._owners, _ptrs=X.ptrs + W_act.ptrs)
            hs_loc = DeviceBlock(ctx, _owners=HS._owners, _ptrs=HX.ptrs + HW_act.ptrs)

        try:
            if incremental_gram and iter_ > 0 and g_xp is not None and iter_ % GRAM_REFRESH != 0:
                # O(R) work for the rows of W only: kernel column order [W_act | X | P_act]
                na = n_act
                P_act, HP_act = P.cols(ind_act), HP.cols(ind_act)
                s_k = DeviceBlock(ctx, _owners=S._owners, _ptrs=W_act.ptrs + X.ptrs + P_act.ptrs)
                hs_k = DeviceBlock(ctx, _owners=HS._owners, _ptrs=HW_act.ptrs + HX.ptrs + HP_act.ptrs)
                gk, tk = gram_pair_top(s_k, hs_k, na)
                ss = np.empty((n_loc, n_loc), dtype=np.complex128)
                shs = np.empty_like(ss)
                ix, iw, ip = slice(0, m), slice(m, m + na), slice(m + na, n_loc)
                pa = m + ind_act                                   # positions of the active P columns in [X | P]
                for dst, src, old in ((ss, gk, g_xp), (shs, tk, t_xp)):
                    dst[ix, ix] = old[:m, :m]
                    dst[ix, ip] = old[:m, pa]
                    dst[ip, ix] = old[pa, :m]
                    dst[ip, ip] = old[np.ix_(pa, pa)]
                    dst[iw, iw] = src[:na, :na]
                    dst[iw, ix] = src[:na, na:na + m]
                    dst[iw, ip] = src[:na, na + m:]
                    dst[ix, iw] = dst[iw, ix].conj().T
                    dst[ip, iw] = dst[iw, ip].conj().T
                ss, shs = hermitize(ss), hermitize(shs)
            else:
                ss, shs = gram_pair(s_loc, hs_loc)
            lambdas, eigvec = rr_small(ss, shs, min_rank=m)
        except np.linalg.LinAlgError:
            return None, None, None
        if np.isnan(lambdas).any() or np.isnan(eigvec).any():
            say(f"{RED}Nan occurs after Rayleigh-Ritz procedure.{RESET}")
            return None, None, None
        lambdas, eigvec = lambdas[:m], np.ascontiguousarray(eigvec[:, :m])
        # _sep_update_after_rr (lobpcg.py:1248-1270) in one pass (+ the next residual, norms and K_P^-1 on the fused path, enqueued
        # before the host-side bookkeeping below so that the two overlap)
        wait_nrms = None
        if fuse_resid:
            wait_nrms = op.update_resid_start(m, n_loc, s_loc, hs_loc, P, HP, eigvec, lambdas, W)
        else:
            L.check(L.lib().pcb_update(ctx.h, m, n_loc, L.ptr_array(s_loc.ptrs), L.ptr_array(hs_loc.ptrs),
                                       L.ptr_array(P.ptrs), L.ptr_array(HP.ptrs), eigvec.ctypes.data), "pcb_update")
        if incremental_gram:
            # Gram pair of the rotated blocks X' = S E, P' = S E_p (E_p = E with the X rows zeroed), all m columns of P'
            e_p = eigvec.copy()
            e_p[:m] = 0.0
            ee = np.concatenate((eigvec, e_p), axis=1)
            g_xp = hermitize(ee.conj().T @ ss @ ee)
            t_xp = hermitize(ee.conj().T @ shs @ ee)

        if wait_nrms is not None:
            next_nrms = wait_nrms()
        say(f"Runtime = {time.time() - t_iter_h:<6.4f}s.")

    ctx.sync()
    t_tot = time.time() - t_tot_h
    say(f"\nA complete procedure of lobpcg is done, {t_tot:<6.2f}s elapsed.")
    info = np.array([iter_, t_tot])
    if history:
        info = np.append(info, res_his[1:iter_])
    x = X.copy()      # detach the result from the 3m-column work block
    return lambdas[:m] - shift, x, info


def lobpcg_sep_softlock_mixedprecision(h_func, p_func, x0, nev, tol=TOL, maxiter=MAXITER, history=False, longortho=False,
                                       trace=None):
    """Soft-locking LOBPCG whose preconditioner input is handed over in single precision (lobpcg.py:494-629): the residual
    block is rounded to complex64 (and widened again) before K_P^-1 -- fused into pcb_residual (precond = 2) -- while H,
    the Gram pair, Rayleigh-Ritz and the update stay complex128.  Differences from ``lobpcg_sep_softlock`` kept from the
    reference: initial lambda = eigvalsh(herm(X^H HX)) (:524), no stagnation rules, NaN residuals raise ValueError (:549),
    no shift argument, and the return value is ``(lambdas[:nev], x[:, :nev], info)`` (:629)."""
    if longortho:
        raise NotImplementedError("longortho (rayleigh_ritz_qr_sep) is outside the ported hot path")
    lam, x, info = lobpcg_sep_softlock(h_func, p_func, x0, nev, shift=0.0, tol=tol, maxiter=maxiter, history=history,
                                       maxstagniter=10 ** 9, trace=trace, _mixed=True)
    if lam is None:
        raise ValueError("Rayleigh-Ritz failed in lobpcg_sep_softlock_mixedprecision")
    return lam[:nev], x[:, :nev], info


def lobpcg_sep_nolock(h_func, p_func, x0, nev, tol=TOL, maxiter=MAXITER, history=False, longortho=False,
                      singleprecision=False):
    """LOBPCG without locking (lobpcg.py:76-193): every column stays in the search space [X W P] until the first `nev`
    residuals are below `tol`.  Same kernels as the soft-locking solver with the full index set as the active set (the
    reference's version only works while all columns are active; this is its intended behaviour).
    Returns (lambdas[:m], x, info); NaN residuals raise ValueError like the reference."""
    lam, x, info = lobpcg_sep_softlock(h_func, p_func, x0, nev, shift=0.0, tol=tol, maxiter=maxiter, history=history,
                                       longortho=longortho, singleprecision=singleprecision, maxstagniter=10 ** 9, _lock=False)
    return lam, x, info


_helpers = {}


def _residual_helper(ctx):
    """An identity-symbol Operator: gives the generic (non-fused) path access to pcb_residual."""
    key = id(ctx)
    if key not in _helpers:
        from .discretization import FourierSymbols
        _helpers[key] = Operator(FourierSymbols(ctx.N, 1, np.eye(3)), 0.0, 0.0, 0.0, None, ctx=ctx)
    return _helpers[key]
