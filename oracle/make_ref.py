"""Recipe for oracle/_ref: a verbatim, git-ignored copy of the reference checkout.

TEST INFRASTRUCTURE ONLY.  /root/reference exists in the authoring container but not on the GPU
box; oracle/_ref/ is listed in .gitignore (never committed: no reference source enters the history)
and NOT in .gpurunignore, so it travels with the snapshot.  With it present, the GPU tests drive the
reference's own recorder (tools/record.py) through the drop-in backend and compare the CUDA step with
the live Numba kernels, and `bench.py --impl reference` times the reference's own CPU path
(nbody/simulation.py:63-317 sequenced as tools/record.py:835-858) instead of the C port.

    python oracle/make_ref.py [--src /root/reference]

Nothing is modified: files are copied byte for byte (Python sources and the docs that name them;
caches, recordings and VCS metadata are skipped).  `__graft_entry__.build()` runs this when the
source tree is present.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SKIP_DIRS = {"__pycache__", ".git", "recordings", "venv", ".venv"}


def make_ref(src: str = "/root/reference", dest: str = DEST) -> int:
    """Copies the reference tree; returns the number of files copied (0 if the source is absent)."""
    if not os.path.isdir(os.path.join(src, "nbody")):
        return 0
    count = 0
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in SKIP_DIRS]
        rel = os.path.relpath(root, src)
        out = os.path.join(dest, rel) if rel != "." else dest
        os.makedirs(out, exist_ok=True)
        for f in files:
            if f.endswith((".pyc", ".pyo")):
                continue
            s, d = os.path.join(root, f), os.path.join(out, f)
            if not os.path.exists(d) or os.path.getsize(d) != os.path.getsize(s) or open(d, "rb").read() != open(s, "rb").read():
                shutil.copyfile(s, d)
            count += 1
    return count


if __name__ == "__main__":
    src = sys.argv[2] if len(sys.argv) > 2 and sys.argv[1] == "--src" else "/root/reference"
    n = make_ref(src)
    print(f"oracle/_ref: {n} files from {src}" if n else f"no reference tree at {src}: nothing copied")
