"""CPU oracle for the cvcs_b200 hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  Nothing under ``cvcs_b200/`` does (tests/test_no_oracle_in_product.py
enforces it).

Two layers:
  * ``oracle.c_oracle``   — ctypes binding of ``cvcs_oracle.c`` (plain C, fp64 softmax arithmetic):
                            the mathematical restatement, independent of torch.
  * ``oracle.torch_path`` — the reference's own call sequence on this path restated call for call
                            with the libraries it uses (torch CPU ``nn.CrossEntropyLoss``,
                            ``torch.max``, a restated torchmetrics update, torchvision ``crop``),
                            i.e. what the reference executes on the host.  Also the timed CPU
                            baseline.
  * ``oracle.ref_shim``   — imports the UNMODIFIED reference modules from /root/reference (build
                            container only) to generate tests/golden/ fixtures.
Parity status: PINNED against golden vectors produced by the reference's own Python and by the
torch/torchvision calls it makes (tests/golden/make_golden.py; tests/test_oracle.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_DIR = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_DIR, "cvcs_oracle.c")
_BUILD = os.path.join(_DIR, "_build")
LIB_PATH = os.path.join(_BUILD, "libcvcs_oracle.so")


def build(force: bool = False) -> str:
    """gcc -O2 -shared the C restatement (strict IEEE: no -ffast-math, no FMA contraction)."""
    os.makedirs(_BUILD, exist_ok=True)
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-o", LIB_PATH, _SRC, "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"building the oracle failed:\n{r.stderr}")
    return LIB_PATH


_lib = None


def clib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib
