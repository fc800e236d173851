"""Batch contract (collate_fn, src/tasks.py:138-158): the oracle restatement and the host mirror against the golden
vectors produced by the reference's own collate_fn (oracle/make_golden_collate.py); the device packing kernel against
both (gpu-marked)."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_port as rp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "collate.npz")
CASES = ["ragged", "single", "empty_caption", "all_empty", "long"]


def load(name):
    z = np.load(GOLD)
    toks, offs = z[name + "/tokens"], z[name + "/offsets"]
    lists = [[int(t) for t in toks[offs[i]:offs[i + 1]]] for i in range(len(offs) - 1)]
    return lists, torch.from_numpy(z[name + "/captions"]), torch.from_numpy(z[name + "/lengths"]), int(z[name + "/max_caption_len"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_collate_matches_reference_golden(name):
    lists, caps, lens, lm = load(name)
    c, l, m = rp.collate_captions(lists)
    assert m == lm and c.dtype == torch.int64 and l.dtype == torch.int32
    assert torch.equal(c, caps) and torch.equal(l, lens)


@pytest.mark.parametrize("name", CASES)
def test_host_collate_fn_matches_reference_golden(name):
    from gic_b200.tasks import collate_fn
    lists, caps, lens, lm = load(name)
    batch = [(torch.full((3, 4, 4), float(i)), toks) for i, toks in enumerate(lists)]
    images, c, l, m = collate_fn(batch)
    assert m == lm and c.dtype == torch.int64 and l.dtype == torch.int32
    assert torch.equal(c, caps) and torch.equal(l, lens)
    assert images.shape == (len(lists), 3, 4, 4) and float(images[-1].max()) == float(len(lists) - 1)


def test_csr_wire_format():
    from gic_b200.tasks import ragged_to_csr
    flat, offs, lm = ragged_to_csr([[5, 6, 7], [], [9]])
    assert flat.tolist() == [5, 6, 7, 9] and offs.tolist() == [0, 3, 3, 4] and lm == 5
    assert flat.dtype == torch.int32 and offs.dtype == torch.int32
    flat, offs, lm = ragged_to_csr([])
    assert flat.numel() == 0 and offs.tolist() == [0] and lm == 2


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_packing_matches_reference_golden(name):
    from gic_b200.tasks import pack_captions_device
    lists, caps, lens, lm = load(name)
    c, l, m = pack_captions_device(lists, "cuda:0")
    torch.cuda.synchronize()
    assert m == lm and c.dtype == torch.int64 and l.dtype == torch.int32
    assert torch.equal(c.cpu(), caps) and torch.equal(l.cpu(), lens)            # integer work: bit-exact


@pytest.mark.gpu
def test_device_packing_large_random_vs_oracle_and_padding():
    from gic_b200.tasks import collate_to_device, pack_captions_device
    rng = np.random.RandomState(7)
    lists = [[int(t) for t in rng.randint(4, 30000, size=n)] for n in rng.randint(0, 60, size=4096)]
    want_c, want_l, want_m = rp.collate_captions(lists)
    c, l, m = pack_captions_device(lists, "cuda:0")
    assert m == want_m and torch.equal(c.cpu(), want_c) and torch.equal(l.cpu(), want_l)
    # a fixed, longer pad length (static shapes for CUDA-graph replay): extra positions are <PAD>
    c2, l2, m2 = pack_captions_device(lists, "cuda:0", max_caption_len=want_m + 7)
    assert m2 == want_m + 7 and torch.equal(c2[:, :want_m].cpu(), want_c) and int(c2[:, want_m:].abs().sum()) == 0
    with pytest.raises(ValueError):
        pack_captions_device(lists, "cuda:0", max_caption_len=want_m - 1)
    batch = [(torch.zeros(3, 2, 2), t) for t in lists[:8]]
    images, c3, l3, m3 = collate_to_device(batch, "cuda:0")
    w3 = rp.collate_captions(lists[:8])
    assert images.is_cuda and torch.equal(c3.cpu(), w3[0]) and torch.equal(l3.cpu(), w3[1]) and m3 == w3[2]
