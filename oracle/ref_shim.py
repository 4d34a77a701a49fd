"""Import the UNMODIFIED reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  The reference's ``utils.py`` / ``dataset.py`` import matplotlib,
seaborn, prettytable and torchmetrics at module level; none is installed here and none is on the
hot path, so they are replaced by inert stand-ins in ``sys.modules`` before the import.  The
reference source itself is executed as-is from its read-only location; nothing is copied.

Used by tests/golden/make_golden.py to produce the committed fixtures.  Not usable on the GPU
box (``/root/reference`` does not exist there) — everything that runs there reads the fixtures.
"""
from __future__ import annotations

import os
import sys
import types

REF_SCRIPTS = "/root/reference/source/scripts"


def available() -> bool:
    return os.path.isdir(REF_SCRIPTS)


class _Anything:
    """Absorbs any attribute access / call (plot helpers that are never reached on the path)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


class _Table:
    """prettytable.PrettyTable stand-in: print_metrics builds tables but returns a dict."""

    def __init__(self, field_names=None):
        self.field_names = field_names
        self.rows = []
        self.align = "r"

    def add_row(self, row):
        self.rows.append(row)

    def __str__(self):
        return "\n".join(str(r) for r in self.rows)


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    def _missing(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Anything()

    m.__getattr__ = _missing  # type: ignore[attr-defined]
    sys.modules[name] = m
    return m


def load():
    """Returns (utils, dataset, converters) — the reference's own modules."""
    if not available():
        raise RuntimeError("/root/reference is not present (GPU box?): use the committed tests/golden fixtures")
    from oracle.torch_path import RestatedConfusionMatrix

    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if name not in sys.modules:
            _stub(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if "prettytable" not in sys.modules:
        _stub("prettytable", PrettyTable=_Table)
    if "torchmetrics" not in sys.modules:
        _stub("torchmetrics")
        _stub("torchmetrics.classification", MulticlassConfusionMatrix=RestatedConfusionMatrix)
        _stub("torchmetrics.segmentation", MeanIoU=_Anything)
    if REF_SCRIPTS not in sys.path:
        sys.path.insert(0, REF_SCRIPTS)
    cwd = os.getcwd()
    try:
        import converters  # noqa: F401  (reference module)
        import dataset     # noqa: F401
        import utils       # noqa: F401
    finally:
        os.chdir(cwd)
    return sys.modules["utils"], sys.modules["dataset"], sys.modules["converters"]
