"""Zero-line swap: with this directory ahead of site-packages, AGPlace's ``import faiss``
(reference test.py:2, datasets/datasets_ws_kitti360.py:4) resolves to the B200 engine."""
from agplace_b200 import FLT_MAX, METRIC_L2, IndexFlatL2  # noqa: F401

__all__ = ["IndexFlatL2", "METRIC_L2"]
