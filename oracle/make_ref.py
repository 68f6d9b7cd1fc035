#!/usr/bin/env python
"""Recipe for oracle/_ref/: a byte-for-byte copy of the UNMODIFIED reference package, made to travel to the GPU box.

TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT.  The reference (amidos2006/gym-pcgrl) is pure Python, so "building"
it means copying its `gym_pcgrl/` package -- only the .py files the hot path imports, no sprites / notebooks / models
-- from /root/reference into the git-ignored oracle/_ref/ together with the gym stand-in the build container needs
(tests/golden/ref_shim.py: gym is not installed anywhere in this image).  Nothing is edited; oracle/_ref/ is listed
in .gitignore (it never enters the history) but not in .gpurunignore, so `bench.py --impl reference` and the
`cpu_baseline` leg can time the reference's own PcgrlEnv.step (gym_pcgrl/envs/pcgrl_env.py:129-150) on the GPU
box's host cores.  /root/reference itself does not exist there.

    python oracle/make_ref.py            # copy if /root/reference is present, else keep whatever is there
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("PCGRL_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
SHIM = os.path.join(HERE, "..", "tests", "golden", "ref_shim.py")


def available():
    return os.path.isfile(os.path.join(DST, "gym_pcgrl", "envs", "pcgrl_env.py")) and os.path.isfile(os.path.join(DST, "ref_shim.py"))


def build(force=False):
    """Returns oracle/_ref if it holds the reference package afterwards, else None."""
    src_pkg = os.path.join(REF_SRC, "gym_pcgrl")
    if not os.path.isdir(src_pkg):
        return DST if available() else None
    if available() and not force:
        same = all(filecmp.cmp(os.path.join(src_pkg, rel), os.path.join(DST, "gym_pcgrl", rel), shallow=False)
                   for rel in _py_files(src_pkg)) and filecmp.cmp(SHIM, os.path.join(DST, "ref_shim.py"), shallow=False)
        if same:
            return DST
    shutil.rmtree(DST, ignore_errors=True)
    for rel in _py_files(src_pkg):
        out = os.path.join(DST, "gym_pcgrl", rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(src_pkg, rel), out)
    shutil.copyfile(SHIM, os.path.join(DST, "ref_shim.py"))
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Unmodified copy of %s/gym_pcgrl (*.py only) + tests/golden/ref_shim.py, made by oracle/make_ref.py.\n"
                "Git-ignored measurement infrastructure: never edit, never commit.\n" % REF_SRC)
    return DST


def _py_files(root):
    out = []
    for d, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                out.append(os.path.relpath(os.path.join(d, f), root))
    return sorted(out)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
