"""TEST / BASELINE INFRASTRUCTURE ONLY -- "install" of the unmodified reference into ``baseline/_ref/``.

The reference is not a Python package (no setup.py / pyproject.toml at its root, only ``sampler/setup.py`` for the
orphan CUDA extension), so ``pip install --target baseline/_ref /root/reference`` has nothing to install; the
equivalent for a plain script tree is a pristine copy.  ``baseline/_ref/`` is git-ignored (never part of the history)
but NOT gpurun-ignored: it travels to the GPU box, where ``/root/reference`` does not exist, so that

  * ``bench.py --impl reference`` times the reference's OWN code (its geometry.py / submodule.py / update.py) on the
    host cores, and
  * ``tests/test_gpu_dropin.py`` can build the real ``continuous_IGEVStereo`` / ``continuous_RaftStereo`` graphs,
    install this library into them and compare final disparities at the BASELINE shapes.

Nothing under ``any-stereo_b200/`` ever imports it.  Run by ``__graft_entry__.build()`` when ``/root/reference`` exists.

    python oracle/install_ref.py            # build container only
"""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("ANYSTEREO_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
KEEP_EXT = (".py", ".cu", ".cpp", ".h", ".yaml", ".md")


def _files(root):
    out = []
    for d, dirs, files in os.walk(root):
        dirs[:] = [x for x in dirs if x not in (".git", "__pycache__")]
        for f in files:
            if f.endswith(KEEP_EXT):
                out.append(os.path.relpath(os.path.join(d, f), root))
    return sorted(out)


def install(verbose=True) -> bool:
    """Copy the reference tree; returns True when baseline/_ref is present and identical to the source."""
    if not os.path.isdir(os.path.join(SRC, "models")):
        return os.path.isdir(os.path.join(DST, "models"))
    files = _files(SRC)
    n_new = 0
    for rel in files:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if os.path.exists(d) and filecmp.cmp(s, d, shallow=False):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        n_new += 1
    with open(os.path.join(DST, "INSTALLED_FROM"), "w") as f:
        f.write("%s (%d files, unmodified copy; see oracle/install_ref.py)\n" % (SRC, len(files)))
    if verbose:
        print("baseline/_ref: %d files (%d copied now)" % (len(files), n_new))
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
