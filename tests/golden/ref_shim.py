"""Import shim for the UNMODIFIED reference: puts the byte-for-byte copy that `tools/install_ref.py` materialises under
`baseline/_ref/` (git-ignored) on `sys.path`, together with the case aliases and the tensordict / matplotlib stubs it needs
on this image.  Used by the golden generators; tests and the bench go through `oracle/ref_live.py`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def install():
    from tools import install_ref
    return install_ref.activate()


if __name__ == "__main__":
    print(install())
