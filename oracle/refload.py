"""Load the UNMODIFIED reference modules on the CPU (test infrastructure only).

Only usable where ``/root/reference`` exists (this container, not the GPU box).
Copies ``paper_2/*.py`` + ``dielectric_examples`` to a scratch directory outside
the repository (the reference writes index/JSON files relative to its cwd,
``environment.py:19-20``, ``dielectric.py:87``), puts ``oracle/refshim`` (the
NumPy-backed ``cupy``/``cupyx`` stand-ins) on ``sys.path`` and imports the
modules from there.  Nothing is copied into the repository.
"""
import importlib
import os
import shutil
import sys

REFERENCE = os.environ.get("PC_REFERENCE", "/root/reference")
SCRATCH = os.environ.get("PC_REF_SCRATCH", "/tmp/pc_reference_scratch")
_HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    return os.path.isdir(os.path.join(REFERENCE, "paper_2"))


def load(modules=("environment", "dielectric", "discretization", "pcfft", "orthogonalization",
                  "lobpcg", "numerical_experiments")):
    """Return dict name -> reference module; chdir()s into the scratch copy."""
    if not available():
        raise RuntimeError("reference tree not present (expected on the GPU box)")
    src = os.path.join(REFERENCE, "paper_2")
    dst = os.path.join(SCRATCH, "paper_2")
    if not os.path.isdir(dst):
        os.makedirs(dst)
        for f in os.listdir(src):
            if f.endswith(".py"):
                shutil.copy(os.path.join(src, f), dst)
        for d in ("dielectric_examples/edge_dofs", "dielectric_examples/volume_dofs",
                  "output/chiral", "output/pseudochiral_trivial", "output/pseudochiral_crossdof"):
            os.makedirs(os.path.join(dst, d), exist_ok=True)
        for sub in ("edge_dofs", "volume_dofs"):
            s = os.path.join(src, "dielectric_examples", sub)
            for f in os.listdir(s):
                shutil.copy(os.path.join(s, f), os.path.join(dst, "dielectric_examples", sub))
    os.chdir(dst)
    for p in (os.path.join(_HERE, "refshim"), dst):
        if p not in sys.path:
            sys.path.insert(0, p)
    return {m: importlib.import_module(m) for m in modules}
