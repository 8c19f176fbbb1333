"""``faiss.contrib`` of the shim: only what the reference imports (anyloc/utilities.py:14)."""
