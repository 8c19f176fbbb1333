"""``import faiss.contrib.torch_utils`` (reference anyloc/utilities.py:14).

In faiss this module monkey-patches the index classes so that ``add`` / ``search`` accept ``torch.Tensor``s.  The
engine's ``IndexFlatL2`` / ``IndexFlatIP`` accept CPU and CUDA tensors natively (agplace_b200/index.py), so importing
this module has nothing left to patch."""
