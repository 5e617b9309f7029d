"""Drop-in import path.  `from distributed_faiss.client import IndexClient` etc. keep working
for code written against facebookresearch/distributed-faiss; everything resolves to the
B200-native implementation in `distributed_faiss_b200`."""
