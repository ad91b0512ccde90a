"""Worker for tests/test_kpath_sharding.py: one rank of a world-size-N gloo job on the CPU (host-emulation build of the
kernels).  Runs numerical_experiments.bandgap_sharded on a 4-point slice of the SC k-path and, on rank 0, writes the
merged JSON; the test compares it with a single-process run of the same indices."""
import importlib
import json
import os
import sys

import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PCB200_QUIET"] = "1"


def main():
    out_dir, lib = sys.argv[1], sys.argv[2]
    td.init_process_group(backend="gloo")
    rank, world = td.get_rank(), td.get_world_size()
    pcb = importlib.import_module("linear-eigenvalue-problems-in-photonic-crystals_b200")
    pcb._lib.use_library(lib)
    pcb.set_device(0)

    def gather(obj):
        out = [None] * world if rank == 0 else None
        td.gather_object(obj, out, dst=0)
        return out

    rows = pcb.numerical_experiments.bandgap_sharded(6, "sc_curv", rank, world, type="chiral", nev=4, seed=1000,
                                                     out_dir=out_dir + "/", indices=[57, 58, 59, 60], gather=gather)
    td.barrier()
    with open(os.path.join(out_dir, f"rows_rank{rank}.json"), "w") as f:
        json.dump({"k": sorted(int(k) for k in rows["iterations"]), "errors": rows["errors"]}, f)
    td.destroy_process_group()


if __name__ == "__main__":
    main()
