"""Input pipeline pieces on either side of the path.

  crop_to_max                 train.py:32-35   the batch cropped to its longest utterance / label sequence
  frequency_mask_augment      tt/utils.py:315-329   } same names, arguments and random-number consumption as the
  time_mask_augment           tt/utils.py:297-312   } reference (numpy uniform, then random.randint, per mask), so a seeded
                                                      run masks exactly the same bins; on a CUDA fp32 batch the masks of one
                                                      call go out as ONE launch (ttx_spec_mask) instead of ten slice fills
  mask_augment                train.py:41-44   both calls (frequency first, like the reference's nesting) in one launch
  LengthBucketSampler         tt/dataset.py:84-106 pads every utterance to the corpus maximum (410 frames / 42 labels) and
                              train.py crops to the batch maximum: with random batches the lattice work B * max T * max U is
                              mostly padding.  This batch sampler groups utterances of similar length (buckets of sorted
                              indices, shuffled per epoch) and deals the batches out to the ranks so that every rank gets the
                              same number of batches with similar lattice sizes.
``install()`` rebinds the two augment functions in ``tt.utils`` (and in ``train`` if it is imported).
"""
import ctypes
import random

import numpy as np
import torch

from . import _lib


def crop_to_max(inputs, inputs_length, targets, targets_length):
    """train.py:32-35."""
    max_inputs_length = int(inputs_length.max())
    max_targets_length = int(targets_length.max())
    return inputs[:, :max_inputs_length, :], inputs_length, targets[:, :max_targets_length], targets_length


def _draw(extent, max_width, mask_num):
    """The reference's draws, in its order: width = int(np.random.uniform(0, max)), start = random.randint(0, extent - width)."""
    out = []
    for _ in range(mask_num):
        w = int(np.random.uniform(low=0.0, high=max_width))
        out.append((random.randint(0, extent - w), w))
    return out


def _apply(inputs, masks):
    """masks: list of (axis, start, width).  One kernel launch for a CUDA float32 tensor whose last dimension is
    contiguous; the reference's slice assignments for everything else (numpy arrays, CPU tensors, other dtypes)."""
    masks = [m for m in masks if m[2] > 0]
    if not masks:
        return inputs
    if (isinstance(inputs, torch.Tensor) and inputs.is_cuda and inputs.dtype == torch.float32 and inputs.dim() == 3 and
            inputs.stride(2) == 1 and inputs.size(0) <= 65535):
        lib = _lib.get()
        dev = inputs.device
        for i in range(0, len(masks), 64):
            part = masks[i:i + 64]
            arr = (ctypes.c_int32 * (3 * len(part)))(*[v for m in part for v in m])
            with torch.cuda.device(dev):
                idx = dev.index if dev.index is not None else torch.cuda.current_device()
                _lib.check(lib.ttx_spec_mask(ctypes.c_void_p(inputs.data_ptr()), inputs.size(0), inputs.size(1),
                                             inputs.size(2), inputs.stride(0), inputs.stride(1), arr, len(part), idx,
                                             torch._C._cuda_getCurrentRawStream(idx)), "ttx_spec_mask")
        return inputs
    for axis, start, width in masks:
        if axis == 1:
            inputs[:, start:start + width, :] = 0
        else:
            inputs[:, :, start:start + width] = 0
    return inputs


def time_mask_augment(inputs, max_mask_time=5, mask_num=10):
    return _apply(inputs, [(1, s, w) for s, w in _draw(inputs.shape[1], max_mask_time, mask_num)])


def frequency_mask_augment(inputs, max_mask_frequency=5, mask_num=10):
    return _apply(inputs, [(2, s, w) for s, w in _draw(inputs.shape[2], max_mask_frequency, mask_num)])


def mask_augment(inputs, max_mask_frequency=5, max_mask_time=5, mask_num=10):
    """time_mask_augment(frequency_mask_augment(inputs, ...), ...) of train.py:41-44 with a single launch."""
    masks = [(2, s, w) for s, w in _draw(inputs.shape[2], max_mask_frequency, mask_num)]
    masks += [(1, s, w) for s, w in _draw(inputs.shape[1], max_mask_time, mask_num)]
    return _apply(inputs, masks)


class LengthBucketSampler(torch.utils.data.Sampler):
    """Batch sampler (pass as ``batch_sampler=``): yields lists of dataset indices.

    lengths      per-utterance sizes used for grouping -- frames, or frames * (labels + 1) for the lattice
    batch_size   utterances per batch (per rank)
    world, rank  every rank iterates the same shuffled batch list and takes every world-th batch starting at its rank;
                 neighbouring batches in that list come from the same length bucket, so the ranks of one step work on
                 lattices of similar size (DDP waits for the slowest rank)
    bucket       batches per bucket: utterances are sorted by length, cut into buckets of bucket * batch_size * world,
                 shuffled inside the bucket, and the buckets' batch groups are shuffled as units
    Every index appears exactly once per epoch over all ranks (the tail is dropped when drop_last, else padded by
    repeating the shortest utterances so that all ranks see the same number of batches)."""

    def __init__(self, lengths, batch_size, world=1, rank=0, bucket=8, seed=0, drop_last=False):
        self.lengths = np.asarray(lengths)
        self.batch_size, self.world, self.rank = int(batch_size), int(world), int(rank)
        self.bucket, self.seed, self.drop_last = int(bucket), int(seed), bool(drop_last)
        self.epoch = 0
        if not (0 <= self.rank < self.world) or self.batch_size < 1 or self.bucket < 1:
            raise ValueError("bad sampler arguments")

    def set_epoch(self, epoch):
        self.epoch = int(epoch)

    def _groups(self):
        rng = np.random.RandomState(self.seed + self.epoch)
        order = np.argsort(self.lengths, kind="stable")
        per_step = self.batch_size * self.world
        n = len(order)
        if self.drop_last:
            order = order[: n - n % per_step]
        elif n % per_step:
            order = np.concatenate([order[: per_step - n % per_step], order])     # repeat the shortest utterances
        groups = []                                   # one group = the `world` batches of one step
        span = per_step * self.bucket
        for b0 in range(0, len(order), span):
            chunk = order[b0: b0 + span].copy()
            rng.shuffle(chunk)
            groups += [chunk[g0: g0 + per_step] for g0 in range(0, len(chunk), per_step)]
        rng.shuffle(groups)
        return groups

    def __iter__(self):
        for g in self._groups():
            yield g[self.rank * self.batch_size: (self.rank + 1) * self.batch_size].tolist()

    def __len__(self):
        n, per_step = len(self.lengths), self.batch_size * self.world
        return n // per_step if self.drop_last else -(-n // per_step)
