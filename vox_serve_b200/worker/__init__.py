"""Worker package with the reference's names (``vox_serve/worker/__init__.py``): ``ModelWorker`` and
``CudaGraphWorker`` are the B200 worker (decode steps always replay CUDA graphs)."""
from .base import CudaGraphWorker, ModelWorker
from .depth import DepthModelWorker

__all__ = ["ModelWorker", "CudaGraphWorker", "DepthModelWorker"]
