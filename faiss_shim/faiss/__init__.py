"""Zero-line swap: with this directory ahead of site-packages, AGPlace's ``import faiss``
(reference test.py:2, datasets/datasets_ws_kitti360.py:4) resolves to the B200 engine."""
from agplace_b200 import (FLT_MAX, METRIC_INNER_PRODUCT, METRIC_L2, IndexFlat, IndexFlatIP, IndexFlatL2,  # noqa: F401
                          Kmeans, StandardGpuResources, index_cpu_to_gpu)

__all__ = ["IndexFlatL2", "IndexFlatIP", "IndexFlat", "METRIC_L2", "METRIC_INNER_PRODUCT", "StandardGpuResources",
           "index_cpu_to_gpu", "Kmeans"]
