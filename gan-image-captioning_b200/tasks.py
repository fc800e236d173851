"""Batch contract of the reference's data layer (src/tasks.py:138-158 ``collate_fn``) -- the step immediately before the
hot path.  The dataset itself (COCO json, PIL, vocabulary building, src/tasks.py:18-136) is
out of scope; what the path consumes is ``(images, captions[B, Lmax] int64, lengths[B] int32, max_caption_len)`` with
``<PAD>=0, <S>=1, <E>=2, <UNK>=3`` (src/tasks.py:42-49).

``collate_fn`` keeps the reference's signature and return value (host tensors).  ``collate_to_device`` is the B200
path: the ragged token lists cross PCIe once as a flat int32 array + offsets (pinned), and ``gic_pack_captions`` builds
the padded int64 captions and the lengths on the device, so neither the padded tensor nor ``F.one_hot`` of it
(src/training.py:158) is ever built on the host."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from . import _lib

PAD, START, END, UNK = 0, 1, 2, 3


def _token_lists(batch) -> List[Sequence[int]]:
    return [b[1] for b in batch]


def ragged_to_csr(token_lists: Sequence[Sequence[int]], pin: bool = False) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """(tokens int32 [sum len], offsets int32 [B+1], max_caption_len = longest + 2)."""
    lens = [len(t) for t in token_lists]
    offsets = torch.zeros(len(lens) + 1, dtype=torch.int32)
    if lens:
        offsets[1:] = torch.tensor(lens, dtype=torch.int64).cumsum(0).to(torch.int32)
    flat = torch.tensor([int(x) for t in token_lists for x in t], dtype=torch.int32)
    max_caption_len = (max(lens) if lens else 0) + 2                   # src/tasks.py:143-147
    if pin and torch.cuda.is_available():
        flat, offsets = flat.pin_memory(), offsets.pin_memory()
    return flat, offsets, max_caption_len


def collate_fn(batch):
    """Reference signature: list of (image[3,S,S], token_list) -> (images, captions, lengths, max_caption_len), all on
    the host (src/tasks.py:138-158)."""
    image_size = batch[0][0].shape[-1]
    images = torch.zeros(len(batch), 3, image_size, image_size)
    for i, (img, _) in enumerate(batch):
        images[i] = img
    captions, lengths, max_caption_len = collate_captions(_token_lists(batch))
    return images, captions, lengths, max_caption_len


def collate_captions(token_lists):
    """The caption half of collate_fn (src/tasks.py:143-156) on the host: captions[B, longest + 2] int64 = <S>, tokens, <E>,
    <PAD>...; lengths[B] int32 = len + 2; max_caption_len."""
    flat, offsets, max_caption_len = ragged_to_csr(token_lists)
    B = len(token_lists)
    captions = torch.zeros(B, max_caption_len, dtype=torch.long)
    lengths = (offsets[1:] - offsets[:-1] + 2).to(torch.int32)
    pos = torch.arange(max_caption_len).unsqueeze(0)                     # [1, Lm]
    ln = (lengths.long() - 2).unsqueeze(1)                              # [B, 1]
    captions[:, 0] = START
    if flat.numel():
        src = (offsets[:-1].long().unsqueeze(1) + pos - 1).clamp(0, flat.numel() - 1)
        tok = flat.long()[src]
        inside = (pos >= 1) & (pos <= ln)
        captions = torch.where(inside, tok, captions)
    captions = torch.where(pos == ln + 1, torch.full_like(captions, END), captions)
    return captions, lengths, max_caption_len


def pack_captions_device(token_lists: Sequence[Sequence[int]], device, max_caption_len: int = None):
    """Ragged token lists -> (captions[B, Lm] int64, lengths[B] int32, Lm) on ``device`` through gic_pack_captions."""
    _lib.require_cuda()
    flat, offsets, lm = ragged_to_csr(token_lists, pin=True)
    if max_caption_len is None:
        max_caption_len = lm
    if max_caption_len < lm:
        raise ValueError("max_caption_len %d is shorter than the longest caption + 2 = %d" % (max_caption_len, lm))
    B = len(token_lists)
    dev = torch.device(device)
    with torch.cuda.device(dev):
        d_flat = flat.to(dev, non_blocking=True) if flat.numel() else torch.zeros(1, dtype=torch.int32, device=dev)
        d_off = offsets.to(dev, non_blocking=True)
        captions = torch.empty(B, max_caption_len, dtype=torch.int64, device=dev)
        lengths = torch.empty(B, dtype=torch.int32, device=dev)
        _lib.check(_lib.lib().gic_pack_captions(_lib.ptr(d_flat), _lib.ptr(d_off), B, int(max_caption_len),
                                                _lib.ptr(captions), _lib.ptr(lengths), _lib.stream()), "gic_pack_captions")
    return captions, lengths, max_caption_len


def collate_to_device(batch, device):
    """collate_fn with the captions packed on the device: (images (device), captions, lengths, max_caption_len)."""
    image_size = batch[0][0].shape[-1]
    images = torch.stack([b[0] for b in batch]).reshape(len(batch), 3, image_size, image_size).to(device, non_blocking=True)
    captions, lengths, lm = pack_captions_device(_token_lists(batch), device)
    return images, captions, lengths, lm
