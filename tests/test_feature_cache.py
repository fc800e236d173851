"""On-disk cache of the frozen CNN trunk's pooled features (src/generator.py:19-22 runs the trunk
under no_grad on every batch).  Host logic only: no kernels run."""
import numpy as np
import pytest
import torch


def _trunk(images):
    """Stand-in for the frozen ResNet + view (src/generator.py:20-22): any deterministic images -> [B, feature_dim] map."""
    return images.flatten(1)[:, :24] * 2.0 + 1.0


def test_build_once_then_serve_batches_and_persist(tmp_path):
    from gic_b200.feature_cache import CachedFeatureLoader, FeatureCache
    ids = ["img%03d" % i for i in range(10)]
    g = torch.Generator().manual_seed(0)
    images = torch.randn(10, 3, 4, 4, generator=g)
    path = str(tmp_path / "feats")
    cache = FeatureCache.create(path, 10, 24, image_ids=ids)
    assert "img003" not in cache
    with pytest.raises(KeyError):
        cache.get(["img003"])                              # never filled: no silent zeros
    calls = []

    def trunk(x):
        calls.append(len(x))
        return _trunk(x)
    batches = [(ids[i:i + 4], images[i:i + 4]) for i in range(0, 10, 4)]
    assert cache.build(batches, trunk=trunk) == 10
    assert cache.build(batches, trunk=trunk) == 0 and sum(calls) == 10     # resumable: the trunk runs once per image
    want = _trunk(images)
    got = cache.get(["img007", "img000", "img007"])
    assert torch.equal(got, want[[7, 0, 7]])
    with pytest.raises(KeyError):
        cache.get(["nope"])
    del cache
    ro = FeatureCache.open(path)                            # persisted: header + data survive a reopen
    assert len(ro) == 10 and "img009" in ro
    assert torch.equal(ro.get(ids, pin=False), want)
    with pytest.raises(IOError):
        ro.put(["img000"], want[:1])
    # batches in the order adv_loop takes them: (pooled, captions) with the collate contract <S> tokens <E> <PAD>...
    toks = [[4 + (i % 5)] * (1 + i % 3) for i in range(10)]
    loader = CachedFeatureLoader(ro, ids, toks, batch_size=4)
    seen = 0
    for pooled, caps in loader:
        b = pooled.shape[0]
        assert torch.equal(pooled, want[seen:seen + b])
        assert caps.dtype == torch.int64 and int(caps[0, 0]) == 1 and caps.shape[1] == max(len(t) for t in toks[seen:seen + b]) + 2
        seen += b
    assert seen == 10 and len(loader) == 3


def test_fp16_file_halves_the_bytes_and_widens_on_read(tmp_path):
    import os
    from gic_b200.feature_cache import FeatureCache
    path = str(tmp_path / "f16")
    c = FeatureCache.create(path, 6, 32, dtype="float16")
    x = torch.randn(6, 32)
    c.put(range(6), x)
    c.flush()
    assert os.path.getsize(path + ".bin") == 6 * 32 * 2
    y = FeatureCache.open(path).get([5, 1], pin=False)
    assert y.dtype == torch.float32 and torch.allclose(y, x[[5, 1]], rtol=1e-3, atol=1e-3)
    with pytest.raises(ValueError):
        c.put([0], torch.zeros(1, 31))


def test_truncated_data_file_is_rejected(tmp_path):
    from gic_b200.feature_cache import FeatureCache
    path = str(tmp_path / "bad")
    FeatureCache.create(path, 4, 8).flush()
    with open(path + ".bin", "r+b") as f:
        f.truncate(17)
    with pytest.raises(ValueError):
        FeatureCache.open(path)
