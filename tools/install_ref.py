#!/usr/bin/env python
"""Materialise the UNMODIFIED reference under `baseline/_ref/` so that it travels to the GPU box.

    python tools/install_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

torch_DE is pure Python with a `hatchling` build backend that this image lacks, so `pip install` cannot build it
(DESIGN.md section 4); its "build" is a file copy.  This script copies `/root/reference/src/torch_DE` byte for byte to
`baseline/_ref/torch_DE` (git-ignored: reference sources never enter the history) and adds what a case-sensitive
filesystem and this image need to import it -- nothing inside the copied files is edited:

  * three aliases for modules the reference imports under another letter case (it was developed on a case-insensitive
    filesystem): `equations -> Equations` (utils/loss.py:7,9), `utils/loss_weighting.py -> Loss_weighting.py`
    (utils/__init__.py:4), `post/plotter.py -> Plotter.py` (post/__init__.py:1);
  * two stub packages under `baseline/_ref/stubs/` for imports this image does not have and the hot path barely touches:
    `matplotlib` (imported at module top of utils/sampling.py:2, never called on the path) and `tensordict` (a dict-backed
    `TensorDict` with the surface utils/data.py:43-54,98-110 uses).

`baseline/_ref/MANIFEST.json` records a sha256 per copied file so tests can tell the copy is unmodified.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_SRC = "/root/reference/src/torch_DE"
DEST = os.path.join(ROOT, "baseline", "_ref")

TENSORDICT_STUB = '''"""Stand-in for tensordict.TensorDict: only the surface torch_DE/utils/data.py:43-54,98-110 uses."""
import torch


class TensorDict(dict):
    def __init__(self, d=None, batch_size=None, lock=False, **kw):
        super().__init__()
        if d:
            for k, v in d.items():
                self[k] = v
        self.batch_size = batch_size

    def to(self, *a, **k):
        return TensorDict({kk: v.to(*a, **k) for kk, v in self.items()}, batch_size=self.batch_size)

    def select(self, *keys):
        return TensorDict({k: self[k] for k in keys}, batch_size=self.batch_size)

    def __getitem__(self, key):
        if isinstance(key, str):
            return dict.__getitem__(self, key)
        return TensorDict({k: v[key] for k, v in self.items()}, batch_size=None)
'''

ALIASES = (("equations", "Equations"), ("utils/loss_weighting.py", "Loss_weighting.py"), ("post/plotter.py", "Plotter.py"))


def _sha(path):
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def install(force: bool = False) -> str:
    pkg = os.path.join(DEST, "torch_DE")
    if os.path.isdir(pkg) and not force and os.path.exists(os.path.join(DEST, "MANIFEST.json")):
        return DEST
    if not os.path.isdir(REFERENCE_SRC):
        raise RuntimeError(f"{REFERENCE_SRC} is not present: the reference can only be materialised in the build container")
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    shutil.copytree(REFERENCE_SRC, pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.ipynb"))
    manifest = {}
    for d, _, files in os.walk(pkg):
        for f in files:
            p = os.path.join(d, f)
            manifest[os.path.relpath(p, pkg)] = _sha(p)
    for alias, target in ALIASES:
        a = os.path.join(pkg, alias)
        if not os.path.lexists(a):
            os.symlink(target, a)
    stubs = os.path.join(DEST, "stubs")
    os.makedirs(os.path.join(stubs, "matplotlib"))
    os.makedirs(os.path.join(stubs, "tensordict"))
    for f in ("matplotlib/__init__.py", "matplotlib/pyplot.py"):
        open(os.path.join(stubs, f), "w").close()
    with open(os.path.join(stubs, "tensordict", "__init__.py"), "w") as fh:
        fh.write(TENSORDICT_STUB)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REFERENCE_SRC, "files": manifest, "aliases": ALIASES}, fh, indent=1, sort_keys=True)
    return DEST


def available() -> bool:
    return os.path.isdir(os.path.join(DEST, "torch_DE")) or os.path.isdir(REFERENCE_SRC)


def activate() -> str:
    """Put the materialised reference (and its stubs) on sys.path; materialise it first when the sources are here."""
    dest = install()
    for p in (os.path.join(dest, "stubs"), dest):
        if p not in sys.path:
            sys.path.insert(0, p)
    return dest


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
