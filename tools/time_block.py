"""Time the LOBPCG block kernels alone (CUDA events on the library's stream): python tools/time_block.py [N] [m]

Prints one JSON line: full Gram pair, incremental Gram (rows of W only) for several active counts, fused update, residual."""
import importlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pcb = importlib.import_module("linear-eigenvalue-problems-in-photonic-crystals_b200")


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    L = pcb._lib
    ctx = pcb.get_context(N)
    S, HS = ctx.random_block(3 * m, 77), ctx.random_block(3 * m, 78)

    def timed(fn, reps=10):
        fn(); ctx.sync(); ctx.timer_start()
        for _ in range(reps):
            fn()
        return ctx.timer_stop() / reps

    out = {"N": N, "m": m, "GRAM_W": os.environ.get("PCB200_GRAM_W")}
    for n_act in (m, m // 2, m // 4):
        nl = m + 2 * n_act
        G = np.empty((nl, nl), dtype=np.complex128); T = np.empty_like(G)
        s, hs = S[:, :nl], HS[:, :nl]
        t = timed(lambda: L.check(L.lib().pcb_gram2(ctx.h, nl, L.ptr_array(s.ptrs), L.ptr_array(hs.ptrs), G.ctypes.data, T.ctypes.data), "gram2"))
        out[f"gram_full_nact{n_act}_ms"] = round(t, 4)
        t = timed(lambda: L.check(L.lib().pcb_gram2_top(ctx.h, nl, n_act, L.ptr_array(s.ptrs), L.ptr_array(hs.ptrs), G.ctypes.data, T.ctypes.data), "gram2_top"))
        out[f"gram_top_nact{n_act}_ms"] = round(t, 4)
        E = np.ascontiguousarray(np.random.default_rng(0).standard_normal((nl, m)) + 0j) / nl
        t = timed(lambda: L.check(L.lib().pcb_update(ctx.h, m, nl, L.ptr_array(s.ptrs), L.ptr_array(hs.ptrs), L.ptr_array(S[:, 2 * m:].ptrs),
                                                     L.ptr_array(HS[:, 2 * m:].ptrs), E.ctypes.data), "update"))
        out[f"update_nact{n_act}_ms"] = round(t, 4)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
