"""GPU tests of the runner layer and of the multi-GPU modes on real devices.

  * bandgap(): checkpointed band structure over a slice of the k-path at N = 48 -- JSON format of the reference
    (numerical_experiments.py:355-357,464-485), warm-start chain, resume, agreement of every row with eigen_1p / the oracle;
  * bandgap_sharded(): two ranks on one GPU (contiguous chunks, merged JSON) == the single-rank rows bit for bit;
  * large-grid mode with NCCL: world = 2 on two B200s when the box has them (skipped on a one-GPU box): same iteration count
    as the one-GPU solve, eigenvalues to 1e-12.
"""
import importlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import PKG, ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture
def gpu(request):
    pkg = importlib.import_module(PKG)
    pkg._lib.use_library(request.getfixturevalue("cuda_lib"))
    assert pkg.backend() == "cuda-sm_100a"
    yield pkg
    pkg.devarray.drop_contexts()
    pkg.lobpcg._helpers.clear()


def test_bandgap_checkpoint_format_and_rows(gpu, oracle, tmp_path):
    ne = gpu.numerical_experiments
    N, d_flag, typ, nev = 48, "sc_curv", "chiral", 6
    path = str(tmp_path / "bandgap_sc_curv.json")
    idx = [18, 19, 20]                       # 19 = X ... a warm-started chain of three neighbouring k-points
    err = ne.bandgap(N, d_flag, type=typ, indices=idx[:2], nev=nev, seed=500, path=path)
    assert err == []
    rec = json.load(open(path))
    key_it, key_fq = f"{d_flag}_{N}_iterations", f"{d_flag}_{N}_frequencies"
    n_k = len(gpu.dielectric.kpath(d_flag))
    assert len(rec[key_it]) == n_k and len(rec[key_fq]) == n_k
    assert rec[key_it][0] == [0, 0] and rec[key_it][20] == [0, 0]           # uncomputed sentinel (reference format)
    assert rec[key_it][18][0] > 0 and rec[key_it][19][0] > 0
    m = nev + round(0.6 * nev)
    assert len(rec[key_fq][18]) == m
    # resume: only the missing row of `only` is computed, existing rows stay bit-identical
    before = json.dumps(rec[key_fq][18])
    ne.bandgap(N, d_flag, type=typ, nev=nev, seed=500, path=path, only=idx)
    rec2 = json.load(open(path))
    assert json.dumps(rec2[key_fq][18]) == before
    assert rec2[key_it][20][0] > 0
    # the warm-started row agrees with an independent cold solve of that k-point (eigen_1p, whose parity with the reference is
    # pinned by the goldens): converged bands, residual tolerance 1e-4 -> frequencies to ~1e-6
    alphas = gpu.dielectric.kpath(d_flag)
    res = ne.eigen_1p(N, d_flag, alphas[19], type=typ, nev=nev, seed=3)
    got = np.array(rec2[key_fq][19][:nev])
    assert np.max(np.abs(got - np.asarray(res["omega_re"])[:nev])) < 2e-5


def test_bandgap_sharded_two_ranks_one_gpu(gpu, tmp_path):
    ne = gpu.numerical_experiments
    out = str(tmp_path) + "/"
    idx = [57, 58, 59, 60]
    rows = {}
    for rank in (0, 1):
        rows[rank] = ne.bandgap_sharded(24, "sc_curv", rank, 2, type="pseudochiral_trivial", nev=4, seed=1000, out_dir=out, indices=idx)
        assert rows[rank]["errors"] == []
    assert sorted(rows[0]["iterations"]) == [57, 58] and sorted(rows[1]["iterations"]) == [59, 60]
    # the chunk of rank 1 computed stand-alone (fresh process state) gives the same rows: no cross-rank state
    gpu.devarray.drop_contexts()
    again = ne.bandgap_sharded(24, "sc_curv", 1, 2, type="pseudochiral_trivial", nev=4, seed=1000, out_dir=str(tmp_path / "b") + "/", indices=idx)
    for k in (59, 60):
        assert again["iterations"][k][0] == rows[1]["iterations"][k][0]
        assert np.array_equal(np.array(again["frequencies"][k]), np.array(rows[1]["frequencies"][k]))


def test_large_grid_world2_nccl(gpu, tmp_path):
    import ctypes as C
    n = C.c_int()
    gpu._lib.check(gpu._lib.lib().pcb_device_count(C.byref(n)), "pcb_device_count")
    if n.value < 2:
        pytest.skip("needs two GPUs (large-grid mode over NCCL)")
    env = dict(os.environ, PCB200_QUIET="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29591", os.path.join(ROOT, "tools", "run_large_grid.py"), "64", "10", "sc_curv", "pseudochiral_trivial", "check"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert res["world"] == 2
    assert res["iterations"] == res["single_gpu"]["iterations"]
    assert res["single_gpu"]["max_rel_eig_diff"] < 1e-12
