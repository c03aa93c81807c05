"""TEST INFRASTRUCTURE — recipe that lets the UNMODIFIED reference travel to the GPU box.

    python oracle/build_ref.py            # also called by __graft_entry__.build()

`/root/reference` exists only in the build container.  bench.py's `cpu_baseline` leg and its
`--impl reference` arm have to time the reference's OWN implementation of the path on the GPU box's
host cores, so this recipe copies the reference's pure-Python package (`skoots/**/*.py`, byte for
byte) into `oracle/_ref/skoots/`.  `oracle/_ref/` is git-ignored (never part of the history, never
product source) but not gpurun-ignored, so it rides along with the snapshot exactly like the built
`.so`.  `oracle/ref_shim.py` imports the package from `/root/reference` when that exists and from
`oracle/_ref` otherwise.  Nothing under `skoots_b200/` reads it.
"""
import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.environ.get("SKOOTS_REFERENCE_SOURCE", "/root/reference")
TARGET = os.path.join(HERE, "_ref")


def _tree_digest(root: str) -> str:
    h = hashlib.sha256()
    for base, dirs, files in sorted(os.walk(root)):
        dirs.sort()
        for f in sorted(files):
            if f.endswith(".py"):
                p = os.path.join(base, f)
                h.update(os.path.relpath(p, root).encode())
                with open(p, "rb") as fh:
                    h.update(fh.read())
    return h.hexdigest()


def build(verbose: bool = False) -> str:
    """Returns 'copied', 'up to date', or 'absent' (no reference tree here: the prebuilt copy, if any, is kept)."""
    src = os.path.join(SOURCE, "skoots")
    if not os.path.isdir(src):
        return "absent"
    dst = os.path.join(TARGET, "skoots")
    if os.path.isdir(dst) and _tree_digest(dst) == _tree_digest(src):
        return "up to date"
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=lambda d, names: [n for n in names
                                                        if n == "__pycache__" or (os.path.isfile(os.path.join(d, n)) and not n.endswith(".py"))])
    with open(os.path.join(TARGET, "README"), "w") as fh:
        fh.write("Unmodified copy of /root/reference/skoots (*.py), made by oracle/build_ref.py.\n"
                 "Git-ignored test infrastructure: the CPU baseline bench.py times on the GPU box.\n"
                 f"sha256 of the tree: {_tree_digest(dst)}\n")
    if verbose:
        print(f"copied {src} -> {dst}")
    return "copied"


if __name__ == "__main__":
    print(build(verbose=True))
