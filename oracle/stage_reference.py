"""TEST / BASELINE INFRASTRUCTURE ONLY — stages the reference's own Python files into ``baseline/_ref/``.

``/root/reference`` exists only in the build container; the GPU box receives the repo snapshot.  ``baseline/_ref/`` is
git-ignored (the reference's sources never enter this repo's history) but NOT gpurun-ignored, so the staged copy travels to
the box and ``bench.py --impl reference`` / ``cpu_baseline`` can time the UNMODIFIED reference there (BASELINE.md §4,
SURVEY.md §8c last row).  Only the Python / YAML files of the hot path's import closure are copied (no notebooks, fonts,
plots): ``__init__.py``, ``src/*.py``, ``src/env/**``.

Run by ``__graft_entry__.build()`` whenever ``/root/reference`` is present:   python oracle/stage_reference.py
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.environ.get("DRPO_REF_SOURCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
KEEP_EXT = (".py", ".yaml", ".yml", ".json")
SKIP_REL = {"src/viz_cartpole", "src/viz_quadrotor", "src/viz_tracking", "src/resources", "src/offline"}   # plots, fonts, offline tooling


def stage(verbose=False):
    if not os.path.isdir(os.path.join(SRC, "src")):
        return None
    n = 0
    for base, dirs, files in os.walk(SRC):
        rel = os.path.relpath(base, SRC)
        dirs[:] = [d for d in dirs if d != "__pycache__" and not d.startswith(".") and os.path.normpath(os.path.join(rel, d)) not in SKIP_REL]
        if rel != "." and not (rel == "src" or rel.startswith("src" + os.sep) or rel == "config"):
            continue
        for f in files:
            if not f.endswith(KEEP_EXT):
                continue
            out = os.path.join(DST, rel, f)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(os.path.join(base, f), out)
            n += 1
    if verbose:
        print(f"staged {n} reference files into {DST}")
    return DST


if __name__ == "__main__":
    if stage(verbose=True) is None:
        print(f"no reference under {SRC}: nothing staged", file=sys.stderr)
