"""Feature ingest for the decoder hot path: a bf16, fixed-N, pinned feature cache.

The reference reads variable-length fp32 region features from HDF5 per image
(updown-baseline/updown/data/readers.py:21-139: `features[index].reshape(num_boxes, 2048)`), pads a batch with zero
rows to its largest box count in numpy (`_collate_image_features`, updown-baseline/updown/data/datasets.py:623-632)
and copies 295 KB of fp32 per image to the device every iteration (updown-baseline/updown/utils/common.py:20-27). The
kernels round the features to bf16 as their first step (`image_prep`), so a cache that already holds bf16 rows,
zero-padded to a FIXED box count, in pinned host memory moves half the bytes and gives bit-identical results: the
padding-box mask is `sum_f |x| > 0` (updown_cell.py:263) and a zero row stays a zero row.

Everything here is host-side layout work (torch on the CPU); the consumer is `UpDownCaptioner.forward`, which accepts
bf16 `image_features` as they are.
"""
from typing import Dict, Iterable, Optional, Sequence

import torch


def collate_features(image_features_list: Sequence, num_boxes: Optional[int] = None, dtype=torch.bfloat16,
                     pin: bool = False) -> torch.Tensor:
    """`_collate_image_features` (datasets.py:623-632) with a fixed box count and a bf16 result.

    image_features_list: per-image (n_i, F) arrays / tensors (what ImageFeaturesReader.__getitem__ returns).
    num_boxes: rows per image in the result (default: the largest n_i of the list, like the reference); images with
    more boxes are an error - the reference never truncates."""
    feats = [torch.as_tensor(x) for x in image_features_list]
    if not feats:
        raise ValueError("empty feature list")
    F = feats[0].shape[-1]
    n_max = max(int(x.shape[0]) for x in feats)
    N = n_max if num_boxes is None else int(num_boxes)
    if n_max > N:
        raise ValueError(f"an image has {n_max} boxes, more than num_boxes={N}")
    out = torch.zeros(len(feats), N, F, dtype=dtype)
    for i, x in enumerate(feats):
        if x.shape[-1] != F or x.dim() != 2:
            raise ValueError(f"image {i}: expected (n, {F}) features, got {tuple(x.shape)}")
        out[i, : x.shape[0]] = x.to(dtype)
    return out.pin_memory() if pin else out


def pack_features(image_features: torch.Tensor, pin: bool = False) -> torch.Tensor:
    """(B, N, F) fp32 batch -> the bf16 form the kernels consume directly (round to nearest even, zero rows kept)."""
    out = image_features.detach().to("cpu" if not image_features.is_cuda else image_features.device).to(torch.bfloat16).contiguous()
    return out.pin_memory() if (pin and not out.is_cuda) else out


class FeatureCache:
    """image_id -> bf16 (num_boxes, F) rows, held in ONE pinned host buffer with a fixed row count per image, so that a
    batch is a single gather + one host->device copy of B*N*F*2 bytes (147 KB per image at 36 x 2048, against the
    reference's 295 KB of fp32 plus a numpy pad per batch)."""

    def __init__(self, num_boxes: int, feature_size: int, capacity: int, pin: bool = True):
        self.num_boxes, self.feature_size = int(num_boxes), int(feature_size)
        self._buf = torch.zeros(int(capacity), self.num_boxes, self.feature_size, dtype=torch.bfloat16)
        if pin:
            self._buf = self._buf.pin_memory()
        self._slot: Dict[object, int] = {}

    def __len__(self):
        return len(self._slot)

    def __contains__(self, image_id):
        return image_id in self._slot

    def put(self, image_id, features) -> None:
        """features: (n, F) fp32 / bf16 with n <= num_boxes (the reader's per-image array)."""
        x = torch.as_tensor(features)
        if x.dim() != 2 or x.shape[1] != self.feature_size or x.shape[0] > self.num_boxes:
            raise ValueError(f"expected (n <= {self.num_boxes}, {self.feature_size}) features, got {tuple(x.shape)}")
        slot = self._slot.get(image_id)
        if slot is None:
            slot = len(self._slot)
            if slot >= self._buf.shape[0]:
                raise RuntimeError("feature cache is full")
            self._slot[image_id] = slot
        self._buf[slot].zero_()
        self._buf[slot, : x.shape[0]] = x.to(torch.bfloat16)

    def batch(self, image_ids: Iterable, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(B, num_boxes, F) bf16 batch of the given images; `out` may be a pinned staging buffer or a device tensor."""
        idx = torch.tensor([self._slot[i] for i in image_ids], dtype=torch.long)
        if out is None:
            return self._buf.index_select(0, idx)
        if out.is_cuda:
            out.copy_(self._buf.index_select(0, idx).pin_memory(), non_blocking=True)
        else:
            torch.index_select(self._buf, 0, idx, out=out)
        return out
