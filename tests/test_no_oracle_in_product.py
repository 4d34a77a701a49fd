"""The product package must never import, call or link the oracle (or any CPU fallback)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "cvcs_b200")
    offenders = []
    for d, _, files in os.walk(pkg):
        if os.path.basename(d) in ("build", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(d, f)).read()
                if re.search(r"\b(import|from)\s+oracle\b|oracle/|cvcs_oracle|/root/reference", text):
                    offenders.append(os.path.join(d, f))
    assert not offenders, offenders


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    import importlib
    from cvcs_b200 import _lib
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _lib._load()
    except ImportError as e:
        assert "no CPU or PyTorch fallback" in str(e)
    else:
        raise AssertionError("loading a missing library did not raise")


def test_cpu_tensors_are_rejected():
    import torch
    from cvcs_b200 import ops
    try:
        ops.argmax(torch.zeros(1, 3, 4, 4))
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("CPU tensor accepted")
