"""k-path sharding: world-size-2 gloo job on the CPU vs a single-process run.  Each rank solves a contiguous
chunk of k-points (warm-start chain inside the chunk, seeded random start at the chunk head); no data-path collective."""
import importlib
import json
import os
import subprocess
import sys

import numpy as np

from conftest import PKG, ROOT


def test_kpath_chunks_cover_path():
    ne = importlib.import_module(PKG).numerical_experiments
    for n_k, world in ((80, 8), (120, 8), (160, 8), (7, 2), (3, 4), (120, 1)):
        chunks = ne.kpath_chunks(n_k, world)
        assert len(chunks) == world
        flat = [i for c in chunks for i in c]
        assert flat == list(range(n_k))                       # contiguous, ordered, complete
        assert max(len(c) for c in chunks) - min(len(c) for c in chunks) <= 1


def test_bandgap_sharded_world2_matches_single(emu_lib, tmp_path):
    out2, out1 = tmp_path / "w2", tmp_path / "w1"
    out2.mkdir(); out1.mkdir()
    env = dict(os.environ, PCB200_QUIET="1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(ROOT, "tests", "dist_worker.py"), str(out2), emu_lib]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    merged = json.load(open(out2 / "bandgap_sc_curv.json"))
    k0 = json.load(open(out2 / "rows_rank0.json"))["k"]
    k1 = json.load(open(out2 / "rows_rank1.json"))["k"]
    assert k0 == [57, 58] and k1 == [59, 60]

    # single process, same chunking semantics: two independent chains (57,58) and (59,60) with the same seeds
    pcb = importlib.import_module(PKG)
    pcb._lib.use_library(emu_lib)
    ne = pcb.numerical_experiments
    for rank in (0, 1):
        ne.bandgap_sharded(6, "sc_curv", rank, 2, type="chiral", nev=4, seed=1000, out_dir=str(out1) + "/", indices=[57, 58, 59, 60])
    fq2 = merged["sc_curv_6_frequencies"]
    it2 = merged["sc_curv_6_iterations"]
    for rank, ks in ((0, [57, 58]), (1, [59, 60])):
        part = json.load(open(out1 / f"bandgap_sc_curv.rank{rank}.json"))
        for k in ks:
            assert it2[k][0] == part["sc_curv_6_iterations"][k][0]                     # same iteration count
            assert np.array_equal(np.array(fq2[k]), np.array(part["sc_curv_6_frequencies"][k]))   # bitwise (same kernels, same x0)
    # rows outside the requested indices stay "uncomputed" ([0, 0]) as in the reference's checkpoint format
    assert it2[0] == [0, 0] and it2[61] == [0, 0]
    assert all(v > 0 for v in fq2[59][:4])
