"""Large-grid mode (BASELINE config 5): one eigenproblem spread over the GPUs of a node.

The dense LOBPCG phase is ROW-sharded: rank g owns the i2 planes [zb[g], zb[g+1]) of every column of S and HS (a "slab"
context), so residual / preconditioner / update are local and the Gram pair needs exactly one NCCL all-reduce of the
n_loc x n_loc matrices (done inside pcb_gram2).  The operator needs whole columns (3-D FFTs): the active columns are dealt
round-robin to the ranks, gathered from the slabs with one grouped ncclSend/ncclRecv (pcb_slab_exchange), transformed by
the ordinary single-GPU kernels, and scattered back.  The reference has no multi-GPU path; the solver code is the same
`lobpcg_sep_softlock` -- `ShardedOperator` offers the interface of `pcfft.Operator`.
"""
import ctypes as C
import os

import numpy as np

from . import _lib as L
from . import devarray
from .devarray import Context, DeviceBlock
from .pcfft import Operator, OperatorCallable, _sym_consistency


def slab_bounds(N, world):
    """i2-plane boundaries zb[0..world]: contiguous slabs, sizes differing by at most one plane."""
    base, extra = divmod(N, world)
    zb = [0]
    for g in range(world):
        zb.append(zb[-1] + base + (1 if g < extra else 0))
    return zb


class SlabComm:
    """Rank-local state of the large-grid mode: full context (operator), slab context (dense phase), communicator."""

    def __init__(self, N, rank, world, unique_id=None, device=None, host_callbacks=None):
        """unique_id: the 128 bytes of pcb_comm_unique_id() from rank 0, distributed by the caller (e.g. with
        torch.distributed.broadcast_object_list).  host_callbacks=(allreduce, p2p) is accepted by the host-emulation test
        build only."""
        self.N, self.rank, self.world = int(N), int(rank), int(world)
        self.zb = slab_bounds(self.N, self.world)
        if self.zb[self.rank + 1] == self.zb[self.rank]:
            raise ValueError("more ranks than grid planes")
        self.full = devarray.get_context(N, device)
        self.slab = Context(N, self.full.device, slab=(self.zb[rank], self.zb[rank + 1]))
        if unique_id is None and self.world > 1:
            raise ValueError("SlabComm: world > 1 needs the unique_id of rank 0 (new_unique_id(), broadcast by the caller); "
                             "ranks that each generate their own id would block in ncclCommInitRank forever")
        uid = unique_id if unique_id is not None else new_unique_id()
        buf = (C.c_char * 128).from_buffer_copy(bytes(uid))
        L.check(L.lib().pcb_comm_init(self.slab.h, buf, self.rank, self.world), "pcb_comm_init")
        if host_callbacks is not None:
            self._cbs = host_callbacks    # keep the ctypes callbacks alive
            L.check(L.lib().pcb_comm_set_host_callbacks(self.slab.h, C.cast(host_callbacks[0], C.c_void_p),
                                                        C.cast(host_callbacks[1], C.c_void_p)), "pcb_comm_set_host_callbacks")
        self._zb_c = (C.c_int * (self.world + 1))(*self.zb)
        self._work = {}
        # Peer-memory mode (CUDA build, world <= 8): the solver's work blocks S / HS (Context.work_block) are mapped into every
        # rank with CUDA IPC when they are created -- a collective step, reached by all ranks at the same point of the SPMD
        # solver -- so that the operator can read and write the slabs in place over NVLink (ShardedOperator.apply_into).
        self._shared = {}       # allocation base pointer -> (list of peer base pointers, allocation bytes)
        # PCB200_LG_P2P: 0 = NCCL exchange both ways (pipelined over column chunks); 1 = both x passes of the operator read / write
        # the slabs of all ranks over peer memory; 2 = NCCL gather + scatter fused into the last pass over peer memory.
        # Unset = the measured best: 1 on two GPUs (N = 120, 16 columns: 1.94 ms per apply vs 2.82 (mode 2) and 3.86 (mode 0)),
        # 0 on more (N = 256, 32 columns on 8 GPUs: 15.5 ms (mode 0) vs 21.2 (mode 1) and 20.0 (mode 2): with 7/8 of the data remote
        # the x passes -- 16-byte accesses per thread, loads in flight for a fraction of a CTA's life -- reach 280-350 GB/s per GPU
        # over NVLink where NCCL's copy kernels reach 534 GB/s per direction; profiles/README.md)
        mode = os.environ.get("PCB200_LG_P2P")
        self.p2p_mode = int(mode) if mode in ("0", "1", "2") else (1 if self.world <= 2 else 0)
        self.p2p = (host_callbacks is None and L.backend() == "cuda-sm_100a" and self.world <= 8 and self.p2p_mode != 0)
        if self.p2p:
            self.slab._on_work_block = self.share_block
        self.rows_of = [3 * (self.zb[g + 1] - self.zb[g]) * self.N ** 2 for g in range(self.world)]

    def share_block(self, blk):
        """Collective: map the allocation behind `blk` (a slab DeviceBlock) into all ranks."""
        for a in blk._owners:
            if a.ptr in self._shared:
                continue
            peers = (C.c_void_p * self.world)()
            L.check(L.lib().pcb_comm_share(self.slab.h, C.c_void_p(a.ptr), peers), "pcb_comm_share")
            self._shared[a.ptr] = ([int(p or 0) for p in peers], a)

    def peer_pointers(self, blk):
        """For every column of a slab block: its pointer on each rank, or None when the block is not shared."""
        out = []
        stride = 16 * self.slab.R
        for p in blk.ptrs:
            hit = None
            for base, (peers, a) in self._shared.items():
                if any(o is a for o in blk._owners) and base <= p and (p - base) % stride == 0:
                    j = (p - base) // stride
                    hit = [peers[g] + 16 * self.rows_of[g] * j for g in range(self.world)]
                    break
            if hit is None:
                return None
            out.append(hit)
        return out

    def barrier(self):
        """Stream-ordered barrier over the ranks on the slab stream (one-element all-reduce)."""
        L.check(L.lib().pcb_comm_barrier(self.slab.h), "pcb_comm_barrier")

    def exchange(self, to_full, owners, slab_block, full_ptrs):
        own = (C.c_int * len(owners))(*owners)
        fp = (C.c_void_p * len(owners))(*[p if p else None for p in full_ptrs])
        L.check(L.lib().pcb_slab_exchange(self.slab.h, 1 if to_full else 0, len(owners), own, self._zb_c,
                                          L.ptr_array(slab_block.ptrs), fp), "pcb_slab_exchange")

    def work_blocks(self, ncols):
        """Two cached blocks of whole columns on the full context (operator input / output)."""
        have = self._work.get("n", 0)
        if have < ncols:
            self._work = {"n": ncols, "in": self.full.empty(ncols), "out": self.full.empty(ncols)}
        return self._work["in"], self._work["out"]

    def gather_rows(self, slab_block):
        """Whole columns of a slab block on every rank as a host array (tests / small problems): all-gather through the
        exchange, one owner at a time."""
        out = np.zeros((3 * self.N ** 3, slab_block.k), dtype=np.complex128)
        for o in range(self.world):
            win, _ = self.work_blocks(slab_block.k)
            owners = [o] * slab_block.k
            self.exchange(True, owners, slab_block, win.ptrs[:slab_block.k] if o == self.rank else [0] * slab_block.k)
            self.slab.sync()
            if o == self.rank:
                out = win.cols(range(slab_block.k)).get()
        return out

    def close(self):
        self.slab.sync()
        self.full.sync()
        for base, (peers, a) in list(self._shared.items()):
            arr = (C.c_void_p * self.world)(*[p if p else None for p in peers])
            L.check(L.lib().pcb_comm_unshare(self.slab.h, arr), "pcb_comm_unshare")
        self._shared.clear()
        self.slab._on_work_block = None
        self.slab.__dict__.pop("_work_blocks", None)      # shared work space must not outlive its mappings
        L.check(L.lib().pcb_comm_destroy(self.slab.h), "pcb_comm_destroy")


def new_unique_id():
    buf = (C.c_char * 128)()
    L.check(L.lib().pcb_comm_unique_id(buf), "pcb_comm_unique_id")
    return bytes(buf)


LG_CHUNKS = int(os.environ.get("PCB200_LG_CHUNKS", "2"))      # pipeline depth of ShardedOperator.apply_into (1 = no overlap)


class ShardedOperator:
    """pcfft.Operator interface over slab blocks: P and the residual are slab-local, A / H go through whole columns."""

    def __init__(self, comm, a_fft, gamma, shift, pshift, diel=None):
        self.comm = comm
        self.ctx = comm.slab
        self.full = Operator(a_fft, gamma, shift, pshift, diel, ctx=comm.full)
        self.local = Operator(a_fft, gamma, shift, pshift, None, ctx=comm.slab)     # symbols only: residual + preconditioner
        self.gamma, self.shift, self.pshift = self.full.gamma, self.full.shift, self.full.pshift

    def residual(self, x, hx, w, lambdas, precond=True, single=False):
        return self.local.residual(x, hx, w, lambdas, precond=precond, single=single)

    def update_resid(self, *args):
        return self.local.update_resid(*args)       # row-local on the slabs; the norms are all-reduced inside the C ABI

    def update_resid_start(self, *args):
        return self.local.update_resid_start(*args)

    def apply_into(self, mode, src, dst):
        if src.k == 0:
            return dst
        if mode == L.APPLY_P:
            return self.local.apply_into(mode, src, dst)
        cm = self.comm
        k = src.k
        owners = [j % cm.world for j in range(k)]
        kmax = (k + cm.world - 1) // cm.world                   # most columns any rank owns (slot = j // world)
        win, wout = cm.work_blocks(kmax)
        if cm.p2p and cm.p2p_mode == 2:
            dp = cm.peer_pointers(dst)
            if dp is not None:
                # NCCL gather of the input slabs (pipelined over chunks of columns as below), the operator on whole columns, and
                # the scatter of the result fused into the last FFT pass, which stores every row straight into the slab of the
                # rank that owns its plane (posted writes over NVLink).  The gather orders the ranks before (every rank sends
                # after its own earlier work), one stream-ordered barrier after.
                nch = max(1, min(kmax, LG_CHUNKS, 8))
                bounds = [(i * kmax) // nch for i in range(nch + 1)]
                for ci in range(nch):
                    js = [j for j in range(k) if bounds[ci] <= j // cm.world < bounds[ci + 1]]
                    if not js:
                        continue
                    pin = [win.ptrs[j // cm.world] if owners[j] == cm.rank else 0 for j in js]
                    cm.exchange(True, [owners[j] for j in js], src.cols(js), pin)
                    cm.slab.record(ci)
                    mine = [j for j in js if owners[j] == cm.rank]
                    if mine:
                        nm, w = len(mine), cm.world
                        dsts = (C.c_void_p * (nm * w))(*[dp[j][g] for j in mine for g in range(w)])
                        slots = [j // cm.world for j in mine]
                        cm.full.wait_for(cm.slab, ci)
                        L.check(L.lib().pcb_apply_dist(self.full.h, mode, nm, None, dsts, cm._zb_c, w,
                                                       L.ptr_array([win.ptrs[sl] for sl in slots]),
                                                       L.ptr_array([wout.ptrs[sl] for sl in slots])), "pcb_apply_dist")
                cm.full.record(15)
                cm.slab.wait_for(cm.full, 15)
                cm.barrier()
                return dst
        if cm.p2p and cm.p2p_mode == 1:
            sp, dp = cm.peer_pointers(src), cm.peer_pointers(dst)
            if sp is not None and dp is not None:
                # Peer-memory path: every rank applies the operator to its own columns, reading the input slabs of all ranks in
                # the first FFT pass and writing the output slabs of all ranks in the last one (pcb_apply_dist); the two
                # stream-ordered barriers make the producers' writes visible before and the results visible after.
                mine = [j for j in range(k) if owners[j] == cm.rank]
                cm.barrier()
                cm.slab.record(0)
                if mine:
                    nm, w = len(mine), cm.world
                    srcs = (C.c_void_p * (nm * w))(*[sp[j][g] for j in mine for g in range(w)])
                    dsts = (C.c_void_p * (nm * w))(*[dp[j][g] for j in mine for g in range(w)])
                    cm.full.wait_for(cm.slab, 0)
                    L.check(L.lib().pcb_apply_dist(self.full.h, mode, nm, srcs, dsts, cm._zb_c, w,
                                                   L.ptr_array(win.ptrs[:nm]), L.ptr_array(wout.ptrs[:nm])), "pcb_apply_dist")
                cm.full.record(1)
                cm.slab.wait_for(cm.full, 1)
                cm.barrier()
                return dst
        # Pipeline over chunks of slots: the exchange of chunk c+1 (slab stream, NCCL) overlaps the operator on chunk c (full
        # context's stream), and the way back of chunk c overlaps the operator on chunk c+1.  Ordering is by events between the
        # two streams -- the host never blocks here; the slab-stream kernels that follow are ordered behind the last exchange.
        nch = max(1, min(kmax, LG_CHUNKS, 8))
        bounds = [(i * kmax) // nch for i in range(nch + 1)]
        chunks = []
        for ci in range(nch):
            js = [j for j in range(k) if bounds[ci] <= j // cm.world < bounds[ci + 1]]
            if js:
                chunks.append(js)
        for ci, js in enumerate(chunks):
            pin = [win.ptrs[j // cm.world] if owners[j] == cm.rank else 0 for j in js]
            cm.exchange(True, [owners[j] for j in js], src.cols(js), pin)            # slabs -> whole columns on their owners
            cm.slab.record(ci)
            mine = [j // cm.world for j in js if owners[j] == cm.rank]
            if mine:
                cm.full.wait_for(cm.slab, ci)
                self.full.apply_into(mode, win.cols(mine), wout.cols(mine))
            cm.full.record(ci)
        for ci, js in enumerate(chunks):
            pout = [wout.ptrs[j // cm.world] if owners[j] == cm.rank else 0 for j in js]
            cm.slab.wait_for(cm.full, ci)
            cm.exchange(False, [owners[j] for j in js], dst.cols(js), pout)          # whole columns -> slabs
        return dst

    def apply(self, mode, x, out=None):
        blk, was_host = devarray.as_block(self.ctx, x)
        res = DeviceBlock(self.ctx, blk.k, vec=blk.vec)
        self.apply_into(mode, blk, res)
        return res.get() if was_host else res


def pc_mfd_handle_sharded(comm, a_fft, b_fft, Diels, inv_fft, shift=0.0):
    """(A_func, H_func, P_func) over slab blocks: numerical_experiments.pc_mfd_handle for the large-grid mode.
    `Diels` must live on comm.full (discretization.*_handle(n, d_flag) on this rank's device)."""
    gamma, pshift = _sym_consistency(a_fft, b_fft, inv_fft)
    op = ShardedOperator(comm, a_fft, gamma, shift, pshift if pshift is not None else shift, Diels)
    return OperatorCallable(op, L.APPLY_A), OperatorCallable(op, L.APPLY_H), OperatorCallable(op, L.APPLY_P)
