"""On-disk cache of the frozen CNN trunk's pooled features.

The reference's ``Encoder.forward`` runs the ResNet trunk under ``torch.no_grad()`` on every batch of every epoch
(src/generator.py:19-22) although the trunk is never trained: its output for an image is a constant.  The trunk itself is
outside the hot path; what the path consumes is the pooled feature ``[B, feature_dim]`` that
``Encoder.linear`` + ``Encoder.bn`` project (``gic_encoder_fwd``).  This module stores those features once and serves
them to the step:

* ``FeatureCache.create(path, n_images, feature_dim)`` / ``FeatureCache.open(path)`` -- one memory-mapped matrix
  ``[n_images, feature_dim]`` (fp32, or fp16 to halve the file) plus a small JSON header (shape, dtype, image-id -> row
  map, which rows are filled);
* ``put(image_ids, feats)`` fills rows (any callable may produce them: the reference's own ``resnet`` + ``view``, run
  once, off the path); ``get(image_ids)`` returns a batch in PINNED host memory so that the H2D copy of the step is
  asynchronous; rows that were never filled raise ``KeyError`` instead of feeding zeros to the generator;
* ``CachedFeatureLoader`` yields ``(pooled, captions)`` batches in the order ``GANInstructor.adv_loop`` takes them.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

_MAGIC = "gic_b200.feature_cache.v1"


class FeatureCache:
    def __init__(self, path: str, header: dict, mode: str):
        self.path, self.header = path, header
        self.n, self.dim = int(header["n_images"]), int(header["feature_dim"])
        self.dtype = np.dtype(header["dtype"])
        self._mm = np.memmap(self._data_path(path), dtype=self.dtype, mode=mode, shape=(self.n, self.dim))
        self._rows: Dict[str, int] = {str(k): int(v) for k, v in header["rows"].items()}
        self._filled = np.zeros(self.n, dtype=bool)
        self._filled[np.asarray(header.get("filled", []), dtype=np.int64)] = True
        self._writable = mode != "r"
        self._stage: Optional[torch.Tensor] = None

    # ---- files -------------------------------------------------------------------------------------------
    @staticmethod
    def _data_path(path: str) -> str:
        return path + ".bin"

    @staticmethod
    def _head_path(path: str) -> str:
        return path + ".json"

    @classmethod
    def create(cls, path: str, n_images: int, feature_dim: int, dtype: str = "float32", image_ids: Optional[Sequence] = None):
        if dtype not in ("float32", "float16"):
            raise ValueError("feature cache dtype must be float32 or float16")
        ids = list(range(n_images)) if image_ids is None else list(image_ids)
        if len(ids) != n_images or len(set(map(str, ids))) != n_images:
            raise ValueError("image_ids must be %d distinct ids" % n_images)
        header = {"magic": _MAGIC, "n_images": int(n_images), "feature_dim": int(feature_dim), "dtype": dtype,
                  "rows": {str(k): i for i, k in enumerate(ids)}, "filled": []}
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        np.memmap(cls._data_path(path), dtype=np.dtype(dtype), mode="w+", shape=(n_images, feature_dim)).flush()
        with open(cls._head_path(path), "w") as f:
            json.dump(header, f)
        return cls(path, header, "r+")

    @classmethod
    def open(cls, path: str, writable: bool = False):
        with open(cls._head_path(path)) as f:
            header = json.load(f)
        if header.get("magic") != _MAGIC:
            raise ValueError("%s is not a gic_b200 feature cache" % path)
        want = int(header["n_images"]) * int(header["feature_dim"]) * np.dtype(header["dtype"]).itemsize
        have = os.path.getsize(cls._data_path(path))
        if have != want:
            raise ValueError("%s: data file has %d bytes, header says %d" % (path, have, want))
        return cls(path, header, "r+" if writable else "r")

    def flush(self):
        if not self._writable:
            return
        self._mm.flush()
        self.header["filled"] = np.nonzero(self._filled)[0].tolist()
        tmp = self._head_path(self.path) + ".tmp"
        with open(tmp, "w") as f:
            json.dump(self.header, f)
        os.replace(tmp, self._head_path(self.path))          # the header is replaced atomically: a crash keeps the old one

    # ---- rows --------------------------------------------------------------------------------------------
    def __len__(self):
        return self.n

    def __contains__(self, image_id) -> bool:
        r = self._rows.get(str(image_id))
        return r is not None and bool(self._filled[r])

    def rows_of(self, image_ids: Iterable) -> np.ndarray:
        try:
            return np.fromiter((self._rows[str(k)] for k in image_ids), dtype=np.int64)
        except KeyError as e:
            raise KeyError("image id %s is not in the feature cache" % e) from None

    def put(self, image_ids: Iterable, feats: torch.Tensor):
        if not self._writable:
            raise IOError("feature cache opened read-only")
        rows = self.rows_of(image_ids)
        f = feats.detach().to("cpu", torch.float32).numpy()
        if f.shape != (len(rows), self.dim):
            raise ValueError("features must be [%d, %d], got %r" % (len(rows), self.dim, tuple(f.shape)))
        self._mm[rows] = f.astype(self.dtype, copy=False)
        self._filled[rows] = True

    def get(self, image_ids: Iterable, pin: bool = True) -> torch.Tensor:
        """[len(image_ids), feature_dim] fp32, in pinned host memory (a staging buffer that is reused: copy it to the device
        -- ``.to(device, non_blocking=True)`` -- before asking for the next batch, or pass pin=False for a private tensor)."""
        rows = self.rows_of(image_ids)
        missing = rows[~self._filled[rows]]
        if missing.size:
            raise KeyError("feature cache rows never filled: %s" % missing[:8].tolist())
        n = len(rows)
        if not pin:
            return torch.from_numpy(np.asarray(self._mm[rows], dtype=np.float32).copy())
        if self._stage is None or self._stage.shape[0] < n:
            t = torch.empty(max(n, 1), self.dim, dtype=torch.float32)
            try:
                t = t.pin_memory()
            except RuntimeError:                 # no CUDA runtime (CPU-only tests): pageable staging
                pass
            self._stage = t
        out = self._stage[:n]
        out.numpy()[...] = self._mm[rows]        # gather straight into the staging buffer (fp16 files widen here)
        return out

    def build(self, batches: Iterable[Tuple[Sequence, torch.Tensor]], trunk=None, flush_every: int = 64):
        """Fill the cache from an iterable of (image_ids, x): x = already pooled features, or images when ``trunk`` (any
        callable images -> [B, feature_dim], e.g. the reference's frozen resnet + view, src/generator.py:20-22) is given.
        Runs under no_grad, exactly once per image; batches whose ids are all present are skipped (resumable)."""
        done = 0
        with torch.no_grad():
            for i, (ids, x) in enumerate(batches):
                ids = list(ids)
                if all(k in self for k in ids):
                    continue
                f = trunk(x) if trunk is not None else x
                self.put(ids, f.reshape(len(ids), -1))
                done += len(ids)
                if (i + 1) % flush_every == 0:
                    self.flush()
        self.flush()
        return done


class CachedFeatureLoader:
    """Batches for ``GANInstructor.adv_loop`` / ``pretrain`` from a feature cache and the collated captions: yields
    ``(pooled [B, feature_dim] pinned fp32, captions [B, Lmax] int64)``; the image never has to be decoded or pushed through the
    trunk again (the reference's DataLoader + Encoder.resnet do both every epoch, src/training.py:28-32, src/generator.py:19-22)."""

    def __init__(self, cache: FeatureCache, image_ids: Sequence, token_lists: Sequence[Sequence[int]], batch_size: int,
                 shuffle: bool = False, seed: int = 1008, drop_last: bool = False):
        if len(image_ids) != len(token_lists):
            raise ValueError("one caption per image id")
        self.cache, self.ids, self.toks = cache, list(image_ids), list(token_lists)
        self.bs, self.shuffle, self.seed, self.drop_last, self.epoch = int(batch_size), shuffle, seed, drop_last, 0

    def __len__(self):
        n = len(self.ids)
        return n // self.bs if self.drop_last else (n + self.bs - 1) // self.bs

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        from .tasks import collate_captions
        order = list(range(len(self.ids)))
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            order = torch.randperm(len(order), generator=g).tolist()
        self.epoch += 1
        for b in range(len(self)):
            idx = order[b * self.bs:(b + 1) * self.bs]
            caps, _lengths, _lmax = collate_captions([self.toks[i] for i in idx])
            yield self.cache.get([self.ids[i] for i in idx]), caps
