"""Large-grid mode on real GPUs: python -m torch.distributed.run --nproc-per-node G tools/run_large_grid.py N [nev] [lattice] [type] [check]

One eigenproblem (k = (pi,pi,pi)) spread over G GPUs: row-sharded dense phase + NCCL all-reduce of the Gram pair, operator
on whole columns through the slab exchange.  With `check`, rank 0 also solves the same problem on one GPU (same x0) and
compares eigenvalues / iteration counts.  torch.distributed (gloo) is only used to hand out the NCCL unique id."""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch.distributed as td

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PCB200_QUIET", "1")


def main():
    N = int(sys.argv[1])
    nev = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    d_flag = sys.argv[3] if len(sys.argv) > 3 else "sc_curv"
    typ = sys.argv[4] if len(sys.argv) > 4 else "chiral"
    check = len(sys.argv) > 5 and sys.argv[5] == "check"
    td.init_process_group(backend="gloo")
    rank, world = td.get_rank(), td.get_world_size()
    pcb = importlib.import_module("linear-eigenvalue-problems-in-photonic-crystals_b200")
    pcb.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    sh, mfd, ne = pcb.sharded, pcb.discretization, pcb.numerical_experiments
    uid = [sh.new_unique_id() if rank == 0 else None]
    td.broadcast_object_list(uid, src=0)
    comm = sh.SlabComm(N, rank, world, unique_id=uid[0])
    alpha = np.array([np.pi, np.pi, np.pi])
    m = nev + round(0.6 * nev)
    relax, pnt = mfd.set_relaxation(alpha)
    a_fft, b_fft = mfd.fft_blocks(N, 1, pcb.dielectric.diel_info(d_flag, option="ct"), alpha=alpha)
    inv_fft = mfd.inverse_3_times_3_B(b_fft, pnt, relax[0])
    Diels = getattr(mfd, typ + "_handle")(N, d_flag)
    A, H, P = sh.pc_mfd_handle_sharded(comm, a_fft, (pnt * b_fft[0], pnt * b_fft[1]), Diels, inv_fft, relax[0])
    x0 = comm.slab.random_block(m, 99)
    td.barrier()
    t0 = time.time()
    lam, x, info = pcb.lobpcg.lobpcg_sep_softlock(H, P, x0, nev)
    wall = time.time() - t0
    w_pnt, w_re = ne.recompute_normalize_print(lam[:nev], x[:, :nev], A, relax[0])
    free, total = comm.full.mem_info()
    out = {"N": N, "world": world, "nev": nev, "m": m, "lattice": d_flag, "type": typ, "iterations": int(info[0]), "solver_s": float(info[1]),
           "wall_s": wall, "ms_per_iteration": 1e3 * float(info[1]) / max(1, int(info[0])), "omega_re": [float(v) for v in w_re],
           "gpu_mem_used_GB_rank0": (total - free) / 1e9}
    if check and rank == 0:
        A1, H1, P1 = ne.pc_mfd_handle(a_fft, (pnt * b_fft[0], pnt * b_fft[1]), Diels, inv_fft, relax[0])
        x1 = comm.full.random_block(m, 99)
        lam1, xs1, info1 = pcb.lobpcg.lobpcg_sep_softlock(H1, P1, x1, nev)
        out["single_gpu"] = {"iterations": int(info1[0]), "solver_s": float(info1[1]),
                             "max_rel_eig_diff": float(np.max(np.abs(lam[:nev] - lam1[:nev]) / np.abs(lam1[:nev])))}
    td.barrier()
    if rank == 0:
        print(json.dumps(out), flush=True)
    comm.close()
    td.destroy_process_group()


if __name__ == "__main__":
    main()
