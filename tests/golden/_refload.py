"""Loads the real reference (read-only, /root/reference) with a stub matplotlib.

Only usable in the build container; used by make_golden.py to produce the committed
fixtures.  Nothing in the GPU tests, smoke() or bench.py imports this.
"""
import os
import sys
import tempfile


def load_reference(path="/root/reference"):
    stub = os.path.join(tempfile.gettempdir(), "ali_stub_mpl")
    os.makedirs(os.path.join(stub, "matplotlib"), exist_ok=True)
    open(os.path.join(stub, "matplotlib", "__init__.py"), "w").close()
    with open(os.path.join(stub, "matplotlib", "pyplot.py"), "w") as f:
        f.write("def __getattr__(n):\n    return lambda *a, **k: None\n")
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "nbcache"))
    sys.path[:0] = [stub, path]
    import Anis_TTF_rays as ref  # noqa: E402
    sys.path.remove(stub)
    sys.path.remove(path)
    ref.tqdm_disable = True
    return ref
