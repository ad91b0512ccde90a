import importlib, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pcb = importlib.import_module("linear-eigenvalue-problems-in-photonic-crystals_b200")
ne = pcb.numerical_experiments
os.makedirs("gpurun_out/dbg/chiral", exist_ok=True)
p = "gpurun_out/dbg/chiral/bandgap_fcc.json"
if os.path.exists(p): os.remove(p)
errs = ne.bandgap(int(sys.argv[1]) if len(sys.argv) > 1 else 120, "fcc", type="chiral", indices=[58, 59, 60, 61], seed=1000, path=p)
print("ERRS", errs)
