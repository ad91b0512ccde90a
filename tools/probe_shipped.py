import json, os, sys, importlib, numpy as np
os.environ["PCB200_QUIET"]="1"
sys.path.insert(0, os.getcwd())
pcb = importlib.import_module("linear-eigenvalue-problems-in-photonic-crystals_b200")
rows = json.load(open("tests/golden/shipped_bands.json"))["rows"]
for row in rows:
    alpha = pcb.dielectric.kpath(row["d_flag"])[row["k_index"]]
    try:
        res = pcb.numerical_experiments.eigen_1p(120, row["d_flag"], alpha, type=row["type"], nev=10, seed=7 + row["k_index"])
        want = np.array(row["frequencies"][:10])
        print(row["type"], row["d_flag"], row["k_index"], "iters", int(res["info"][0]), "sec %.3f" % res["info"][1], "ref iters/sec", row["iterations"],
              "maxdiff %.2e" % np.max(np.abs(res["omega_re"] - want)), "maxres %.1e" % res["residuals"].max(), flush=True)
    except Exception as e:
        print(row["type"], row["d_flag"], row["k_index"], "FAILED", repr(e)[:200], flush=True)
