"""Extract a few rows of the reference's SHIPPED band-structure results (paper_2/output/**/bandgap_*.json, computed by the
authors on an RTX 4090 D with tol 1e-4) into tests/golden/shipped_bands.json as known-answer vectors at N = 120.

    python oracle/make_shipped_golden.py        (build container only: reads /root/reference)
"""
import json
import os

REF = "/root/reference/paper_2/output"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# (sub-directory = dielectric type, file, key prefix in the file, d_flag today, eps_opt, k-indices to keep)
PICK = [
    ("chiral", "bandgap_fcc0.json", "fcc", "fcc", 0, [0, 1, 2, 39, 59, 119]),
    ("chiral", "bandgap_sc_curv0.json", "sc_curv", "sc_curv", 0, [19, 59, 79]),
    ("chiral", "bandgap_bcc_single_gyroid0.json", "bcc_single_gyroid", "bcc_sg", 0, [59]),
    ("pseudochiral_trivial", "bandgap_sc_curv0.json", "sc_curv", "sc_curv", 0, [59]),
    ("pseudochiral_crossdof", "bandgap_bcc_sg0.json", "bcc_single_gyroid", "bcc_sg", 0, [59]),
    ("pseudochiral_crossdof", "bandgap_fcc0.json", "fcc", "fcc", 0, [39]),
]
out = []
for typ, fn, prefix, d_flag, eps_opt, idx in PICK:
    lib = json.load(open(os.path.join(REF, typ, fn)))
    fq, it = lib[f"{prefix}_120_frequencies"], lib[f"{prefix}_120_iterations"]
    for i in idx:
        out.append({"type": typ, "file": f"paper_2/output/{typ}/{fn}", "d_flag": d_flag, "eps_opt": eps_opt, "N": 120,
                    "k_index": i, "frequencies": fq[i], "iterations": it[i]})
with open(os.path.join(ROOT, "tests", "golden", "shipped_bands.json"), "w") as f:
    json.dump({"source": "reference repository, shipped results (RTX 4090 D, tol 1e-4)", "rows": out}, f, indent=1)
print(len(out), "rows")
