"""The C-ABI shared library loads and exports every symbol include/agpknn.h declares (CPU only,
no compute calls); without a GPU the product path fails loudly instead of falling back."""
import re
from pathlib import Path

import numpy as np
import pytest

from agplace_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "agpknn.h").read_text()
    return sorted(set(re.findall(r"AGP_API[^;(]*?\b(agp_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for must in ("agp_index_create", "agp_index_add", "agp_index_search", "agp_index_reset", "agp_index_free",
                 "agp_index_ntotal", "agp_merge_topk", "agp_recall_at_n", "agp_last_error"):
        assert must in names
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES and include/agpknn.h must list the same symbols"


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), f"libagpknn.so does not export {name}"
    assert b"sm_100a" in lib.agp_version()


def test_no_cpu_fallback_without_gpu(gpu_available):
    if gpu_available:
        pytest.skip("a GPU is visible; the no-device error path is covered on the CPU box")
    import agplace_b200 as agp
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback|failed"):
        agp.IndexFlatL2(8)


def test_argument_validation_needs_no_gpu():
    import ctypes
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.agp_index_create(0, 0, 0, ctypes.byref(h)) == -1        # AGP_EINVAL: d <= 0
    assert b"d must be positive" in lib.agp_last_error()
    assert lib.agp_index_create(8, 0, 99, ctypes.byref(h)) == -1
    assert lib.agp_index_ntotal(None) == -1
    assert lib.agp_kernel_launches() >= 0


def test_cubin_is_sm100a_with_tcgen05_and_tma():
    """The shipped library carries sm_100a SASS with tensor-core (UTC*MMA), TMEM (LDTM) and TMA ops."""
    import shutil, subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in out, f"{mnemonic} missing from SASS"


def test_faiss_shim_exports_the_names_the_reference_uses():
    """`import faiss` resolves to the engine when faiss_shim/ is ahead on sys.path: every faiss name AGPlace touches
    (test.py:27, datasets_ws_*.py, anyloc/utilities.py:446-453, model/aggregation.py:170) must exist there."""
    import importlib
    import sys
    sys.path.insert(0, str(ROOT / "faiss_shim"))
    try:
        sys.modules.pop("faiss", None)
        faiss = importlib.import_module("faiss")
        for name in ("IndexFlatL2", "IndexFlatIP", "IndexFlat", "METRIC_L2", "METRIC_INNER_PRODUCT", "StandardGpuResources",
                     "index_cpu_to_gpu", "Kmeans"):
            assert hasattr(faiss, name), name
        assert faiss.METRIC_L2 == 1 and faiss.METRIC_INNER_PRODUCT == 0
        importlib.import_module("faiss.contrib.torch_utils")        # reference anyloc/utilities.py:14
    finally:
        sys.path.remove(str(ROOT / "faiss_shim"))
        for m in [m for m in sys.modules if m == "faiss" or m.startswith("faiss.")]:
            sys.modules.pop(m, None)
