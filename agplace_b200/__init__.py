"""agplace_b200 -- B200-native exact L2 descriptor retrieval for AGPlace's hot path.

``import agplace_b200 as faiss`` is a one-line swap at the reference call sites
(test.py:27-32 and the mining helpers in datasets/datasets_ws_*.py): ``faiss.IndexFlatL2`` here is
a hand-written sm_100a CUDA engine behind the C ABI in ``include/agpknn.h``.
"""
from .index import (FLT_MAX, METRIC_INNER_PRODUCT, METRIC_L2, IndexFlat, IndexFlatIP, IndexFlatL2, StandardGpuResources,
                    best_of_lists, default_device, index_cpu_to_gpu, positives_to_csr, radius_neighbors, recall_hits)

__all__ = ["IndexFlatL2", "IndexFlatIP", "IndexFlat", "METRIC_L2", "METRIC_INNER_PRODUCT", "FLT_MAX", "StandardGpuResources",
           "index_cpu_to_gpu", "radius_neighbors", "best_of_lists", "default_device", "positives_to_csr", "recall_hits"]
from .kmeans import Kmeans  # noqa: E402

__all__.append("Kmeans")
__version__ = "0.1.0"
