"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if not fn.endswith(".h"):
            continue
        text = open(os.path.join(ROOT, "include", fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b(cvcs_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("cvcs_ce_fused", "cvcs_label_hist", "cvcs_argmax", "cvcs_confmat", "cvcs_tile_normalize",
                 "cvcs_host_ce_fused", "cvcs_last_error", "cvcs_abi_version"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from cvcs_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/ but not exported by the .so"


def test_python_binding_covers_every_declared_symbol():
    from cvcs_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_no_compute_needed_calls():
    from cvcs_b200 import _lib
    assert _lib.abi_version() == 1
    assert _lib.workspace_bytes() >= 64 * 1024
    try:
        _lib.set_option(99, 1)
    except _lib.CvcsError as e:
        assert e.code == _lib.ERR_INVALID_ARG and "unknown option" in e.message
    else:
        raise AssertionError("bad option accepted")
    _lib.set_option(_lib.OPT_CE_PATH, _lib.CE_PATH_AUTO)


def test_library_has_no_torch_dependency():
    import subprocess
    from cvcs_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    names = [line.split()[0] for line in out.splitlines() if line.strip()]      # library names only: the load
    assert names and not any("torch" in n or "c10" in n for n in names)         # addresses are random hex ("...c10...")
