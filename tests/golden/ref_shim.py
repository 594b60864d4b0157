"""Import shim for the UNMODIFIED reference (`/root/reference/src/torch_DE`).

Only usable in the build container (the GPU box has no /root/reference). It is used by
`make_golden.py` to generate the committed fixtures and by nothing else.

The reference was written on a case-insensitive filesystem, so on Linux it
needs three aliases, and it imports two packages this image lacks (`tensordict`, `matplotlib`) whose
surface on the hot path is tiny.  Nothing is copied into the repo: a symlink farm is created in a
temporary directory.
"""
import os
import sys
import tempfile

REFERENCE_SRC = "/root/reference/src/torch_DE"

_TENSORDICT_STUB = '''
import torch
class TensorDict(dict):
    """Minimal stand-in for tensordict.TensorDict (surface used by torch_DE/utils/data.py:43-54,98-110)."""
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


def _link_tree(src, dst):
    """Mirror `src` into `dst` with real directories and symlinked files."""
    os.makedirs(dst, exist_ok=True)
    for name in os.listdir(src):
        s, d = os.path.join(src, name), os.path.join(dst, name)
        if os.path.isdir(s):
            _link_tree(s, d)
        elif not os.path.exists(d):
            os.symlink(s, d)


def install():
    """Make `import torch_DE` resolve to the unmodified reference sources. Returns the shim dir."""
    if not os.path.isdir(REFERENCE_SRC):
        raise RuntimeError("reference sources not present (this only works in the build container)")
    root = os.path.join(tempfile.gettempdir(), "torch_de_ref_shim")
    pkg = os.path.join(root, "torch_DE")
    _link_tree(REFERENCE_SRC, pkg)
    # case aliases (utils/__init__.py:4, utils/loss.py:7,9, post/__init__.py:1)
    for alias, target in (("equations", "Equations"),
                          ("utils/loss_weighting.py", "Loss_weighting.py"),
                          ("post/plotter.py", "Plotter.py")):
        a = os.path.join(pkg, alias)
        if not os.path.lexists(a):
            os.symlink(target, a)
    stubs = os.path.join(root, "stubs")
    os.makedirs(os.path.join(stubs, "matplotlib"), exist_ok=True)
    os.makedirs(os.path.join(stubs, "tensordict"), exist_ok=True)
    for f in ("matplotlib/__init__.py", "matplotlib/pyplot.py"):
        open(os.path.join(stubs, f), "a").close()
    with open(os.path.join(stubs, "tensordict/__init__.py"), "w") as fh:
        fh.write(_TENSORDICT_STUB)
    for p in (stubs, root):
        if p not in sys.path:
            sys.path.insert(0, p)
    return root


if __name__ == "__main__":
    print(install())
