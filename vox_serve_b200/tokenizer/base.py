"""Per-request streaming-codec state container with the reference's interface (``vox_serve/tokenizer/base.py:8-173``:
``DecoderCache`` with ``cache[index]``, ``copy_from``, ``cat`` and ``to``), which the worker uses to carry a codec's
state between detokenize calls (``requests.py:28-29``; ``cuda_graph_worker.py:1216-1241`` concatenates the per-request
caches, copies them into the graph buffer and slices the result back).

One generic structural map does the work: a cache is a dataclass whose fields are tensors, nested caches, lists /
tuples / dicts of those, or plain values; every operation is "apply f to the aligned tensor leaves".
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Any, Callable, List, Sequence

import torch


def _tree_map(f: Callable[[List[torch.Tensor]], Any], trees: Sequence[Any], where: str = "cache") -> Any:
    """f over the aligned tensor leaves of structurally identical ``trees``; containers are rebuilt, plain values must
    agree across the trees and are passed through."""
    head = trees[0]
    if torch.is_tensor(head):
        if not all(torch.is_tensor(t) for t in trees):
            raise TypeError(f"{where}: tensor / non-tensor mismatch")
        return f(list(trees))
    if isinstance(head, DecoderCache):
        if not all(type(t) is type(head) for t in trees):
            raise TypeError(f"{where}: cannot combine {[type(t).__name__ for t in trees]}")
        return type(head)(**{fl.name: _tree_map(f, [getattr(t, fl.name) for t in trees], f"{where}.{fl.name}")
                             for fl in fields(head)})
    if isinstance(head, (list, tuple)):
        if not all(isinstance(t, type(head)) and len(t) == len(head) for t in trees):
            raise ValueError(f"{where}: sequence fields must have the same type and length")
        return type(head)(_tree_map(f, [t[i] for t in trees], f"{where}[{i}]") for i in range(len(head)))
    if isinstance(head, dict):
        if not all(isinstance(t, dict) and t.keys() == head.keys() for t in trees):
            raise ValueError(f"{where}: dict fields must have the same keys")
        return {k: _tree_map(f, [t[k] for t in trees], f"{where}[{k!r}]") for k in head}
    if all(t is None for t in trees) or all(t == head for t in trees):
        return head
    raise TypeError(f"{where}: plain members differ across caches ({[type(t).__name__ for t in trees]})")


@dataclass
class DecoderCache:
    """Base of the model-specific codec caches (an empty dataclass to inherit from)."""

    def __getitem__(self, index: Any) -> "DecoderCache":
        """The same cache restricted along the batch dimension (``index`` applied to dim 0 of every tensor)."""
        return _tree_map(lambda ts: ts[0][index], [self])

    @torch.no_grad()
    def copy_from(self, src: "DecoderCache") -> None:
        """In-place copy of every tensor of ``src`` into this cache."""
        if type(self) is not type(src):
            raise TypeError(f"Cannot copy from {type(src)} to {type(self)}")
        _tree_map(lambda ts: ts[0].copy_(ts[1]), [self, src])

    @classmethod
    def cat(cls, caches: List["DecoderCache"]) -> "DecoderCache":
        """Concatenate caches of one class along the batch dimension."""
        if not caches:
            raise ValueError("caches must be a non-empty list")
        if not all(isinstance(c, type(caches[0])) for c in caches):
            raise TypeError("All caches must be instances of the same cache class")
        return _tree_map(lambda ts: torch.cat(ts, dim=0), list(caches))

    @torch.no_grad()
    def index_copy_(self, index: torch.Tensor, src: "DecoderCache") -> None:
        """Write the batch rows of ``src`` into rows ``index`` (int64 tensor) of this cache: the inverse of
        ``self[index]``.  The worker keeps ONE cache for all its batch slots and moves the rows of the streams a vocoder
        call serves in and out with these two (no host synchronisation, CUDA-graph capturable)."""
        if type(self) is not type(src):
            raise TypeError(f"Cannot copy from {type(src)} to {type(self)}")
        _tree_map(lambda ts: ts[0].index_copy_(0, index, ts[1]), [self, src])

    @torch.no_grad()
    def zero_rows_(self, index: Any) -> None:
        """Reset the state of the batch rows ``index`` (a stream that starts on a recycled slot)."""
        _tree_map(lambda ts: ts[0][index].zero_() if not isinstance(index, torch.Tensor) else ts[0].index_fill_(0, index, 0), [self])

    def to(self, device) -> "DecoderCache":
        """A copy with every tensor on ``device``."""
        return _tree_map(lambda ts: ts[0].to(device), [self])
