"""Alias of distributed_faiss_b200.index_state (drop-in import path, see distributed_faiss/__init__.py)."""
from distributed_faiss_b200.index_state import *  # noqa: F401,F403
from distributed_faiss_b200 import index_state as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
