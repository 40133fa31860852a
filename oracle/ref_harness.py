"""Import harness for the UNMODIFIED reference (/root/reference) -- test infrastructure only.

Only usable in the build container (the GPU box has no /root/reference).  It is used by
tests/golden/make_golden.py to mint golden vectors and by tests that are skipped when the
reference is absent.  Nothing in the product package imports this file.

The reference imports matplotlib / pygame / skimage / casadi at module import time
(systems/ControlAffineSystem.py:5 -> plots/plot_functions.py:11-15); none are installed here, so
permissive stub modules are registered in sys.modules.  The pickled controller
(systems/sytem_CAR.py:120-124) needs torch.load(weights_only=False) and CWD = reference root.
"""
import os
import sys
import types
import contextlib

REF_ROOT = os.environ.get("DENSITY_PLANNER_REF", "/root/reference")


class _Dummy:
    """Permissive attribute sink; raises AttributeError for dunders so inspect/torch keep working."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Dummy()

    def __iter__(self):
        return iter(())

    def __getitem__(self, k):
        return _Dummy()

    def __setitem__(self, k, v):
        pass


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Dummy()


_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.animation", "matplotlib.patches",
    "matplotlib.path", "matplotlib.widgets", "matplotlib.cm", "matplotlib.lines", "matplotlib.ticker",
    "pygame", "pygame.locals", "casadi",
    "skimage", "skimage.io", "skimage.measure", "skimage.transform", "skimage.color", "skimage.util",
    "skimage.filters", "skimage.draw", "skimage.morphology",
]


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "systems"))


def install():
    """Register stubs, put the reference on sys.path, patch torch.load.  Idempotent."""
    import torch

    for name in _STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = _StubModule(name)
                m.__path__ = []
                sys.modules[name] = m
                if "." in name:
                    parent, child = name.rsplit(".", 1)
                    setattr(sys.modules[parent], child, m)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    ctrl = os.path.join(REF_ROOT, "data", "trained_controller")
    if ctrl not in sys.path:
        sys.path.insert(0, ctrl)
    if not getattr(torch.load, "_dp_patched", False):
        _orig = torch.load

        def _load(*a, **k):
            k.setdefault("weights_only", False)
            return _orig(*a, **k)

        _load._dp_patched = True
        torch.load = _load


@contextlib.contextmanager
def ref_cwd():
    """The reference uses relative paths (sytem_CAR.py:120-123, hyperparams.py:49-52)."""
    old = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        yield
    finally:
        os.chdir(old)


def default_args():
    install()
    import hyperparams

    argv = sys.argv
    sys.argv = ["x"]
    try:
        args = hyperparams.parse_args()
    finally:
        sys.argv = argv
    return args


def make_car(args=None):
    install()
    if args is None:
        args = default_args()
    with ref_cwd():
        from systems.sytem_CAR import Car

        return Car(args), args
