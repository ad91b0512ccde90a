"""Worker for tests/test_large_grid_sharding.py: one rank of a gloo job running the LARGE-GRID mode on the host-emulation
build.  The C ABI's collectives (all-reduce of the Gram pair / norms, slab <-> column exchange) are routed to
torch.distributed through the host-callback hook that exists in the emulation build only; on B200 the same code path calls NCCL."""
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PCB200_QUIET"] = "1"


def main():
    out_dir, lib, N, typ = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    td.init_process_group(backend="gloo")
    rank, world = td.get_rank(), td.get_world_size()
    pcb = importlib.import_module("linear-eigenvalue-problems-in-photonic-crystals_b200")
    pcb._lib.use_library(lib)
    pcb.set_device(0)

    @C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_longlong)
    def allreduce(buf, count):
        a = np.ctypeslib.as_array(buf, shape=(count,))
        t = torch.from_numpy(a)
        td.all_reduce(t)
        return 0

    @C.CFUNCTYPE(C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_longlong))
    def p2p(nops, is_send, peer, ptr, nbytes):
        ops, keep = [], []
        for i in range(nops):
            arr = np.ctypeslib.as_array(C.cast(ptr[i], C.POINTER(C.c_uint8)), shape=(nbytes[i],))
            t = torch.from_numpy(arr)
            keep.append(t)
            ops.append(td.P2POp(td.isend if is_send[i] else td.irecv, t, peer[i]))
        for w in td.batch_isend_irecv(ops):
            w.wait()
        return 0

    sh, mfd, ne = pcb.sharded, pcb.discretization, pcb.numerical_experiments
    comm = sh.SlabComm(N, rank, world, unique_id=bytes(128), host_callbacks=(allreduce, p2p))
    d_flag, alpha, nev, m = "sc_curv", np.array([np.pi, np.pi, np.pi]), 6, 10
    relax, pnt = mfd.set_relaxation(alpha)
    a_fft, b_fft = mfd.fft_blocks(N, 1, pcb.dielectric.diel_info(d_flag, option="ct"), alpha=alpha)
    inv_fft = mfd.inverse_3_times_3_B(b_fft, pnt, relax[0])
    Diels = getattr(mfd, typ + "_handle")(N, d_flag)
    A, H, P = sh.pc_mfd_handle_sharded(comm, a_fft, (pnt * b_fft[0], pnt * b_fft[1]), Diels, inv_fft, relax[0])
    x0 = comm.slab.random_block(m, 4242)
    lam, x, info = pcb.lobpcg.lobpcg_sep_softlock(H, P, x0, nev)
    w_pnt, w_re = ne.recompute_normalize_print(lam[:nev], x[:, :nev], A, relax[0])
    xfull = comm.gather_rows(x[:, :nev])
    td.barrier()
    if rank == 0:
        np.save(os.path.join(out_dir, "x.npy"), xfull)
        with open(os.path.join(out_dir, "result.json"), "w") as f:
            json.dump({"lam": [float(v) for v in lam], "iters": int(info[0]), "w_re": [float(v) for v in w_re],
                       "zb": comm.zb, "world": world}, f)
    comm.close()
    td.destroy_process_group()


if __name__ == "__main__":
    main()
