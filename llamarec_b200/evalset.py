"""Input side of the hot path: the eval splits of the reference's LRU dataloader as
pre-padded device tensors.

The reference builds every eval batch on the host: `LRUValidDataset` / `LRUTestDataset.__getitem__`
(dataloader/lru.py:129-150, 153-180) slice the user's history to the last `max_len` items, left-pad with 0 and wrap
it in a LongTensor; a DataLoader (shuffle=False, dataloader/lru.py:73-99) collates and the trainer copies the batch to
the GPU.  The histories do not change between epochs, so here the whole split is laid out ONCE as an int32
`[U, max_len]` matrix (half the id bytes; `lrb_prepare_sequences` accepts int32 ids) plus an int64 `[U]` label
vector, and iteration hands out views -- no workers, no collation, no per-batch H2D.

Iterating yields `(seqs [b, L], labels [b, 1])` in the reference loader's order (users sorted, last batch short), so
`LRURetriever(args, model, val_loader=DeviceEvalSet(...), test_loader=DeviceEvalSet(...))` is a drop-in.
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch


class DeviceEvalSet:
    def __init__(self, u2seq: Mapping[int, Sequence[int]], u2answer: Mapping[int, Sequence[int]], max_len: int,
                 batch_size: int, u2val: Optional[Mapping[int, Sequence[int]]] = None, device="cuda",
                 subset_users: Optional[Sequence[int]] = None):
        """u2val=None: validation split (history = train sequence, dataloader/lru.py:129-150);
        u2val given: test split (history = train + val, users need both a val and a test item, :153-180)."""
        users = sorted(u2seq.keys())
        if u2val is None:
            users = [u for u in users if len(u2answer[u]) > 0]
        else:
            users = [u for u in users if len(u2val[u]) > 0 and len(u2answer[u]) > 0]
        if subset_users is not None:
            users = list(subset_users)
        self.users: List[int] = users
        self.max_len = int(max_len)
        self.batch_size = int(batch_size)
        ids = np.zeros((len(users), self.max_len), dtype=np.int32)
        labels = np.zeros(len(users), dtype=np.int64)
        for row, u in enumerate(users):
            seq = list(u2seq[u]) + (list(u2val[u]) if u2val is not None else [])
            seq = seq[-self.max_len:]
            if seq:
                ids[row, self.max_len - len(seq):] = seq          # left pad with 0 (:147-149, :177-179)
            answer = u2answer[u]
            if len(answer) != 1:
                raise ValueError(f"user {u}: the retrieval path expects one held-out item, got {len(answer)}")
            labels[row] = answer[0]
        self.seqs = torch.from_numpy(ids).to(device)
        self.labels = torch.from_numpy(labels).to(device)

    @classmethod
    def from_tensors(cls, seqs: torch.Tensor, labels: torch.Tensor, batch_size: int, device="cuda") -> "DeviceEvalSet":
        """An eval split that already exists as a left-padded id matrix [U, L] (+ labels [U]): stored as int32 on
        the device and iterated like the reference's DataLoader (shuffle=False, last batch short)."""
        self = cls.__new__(cls)
        self.users = list(range(1, seqs.shape[0] + 1))
        self.max_len = int(seqs.shape[1])
        self.batch_size = int(batch_size)
        if int(seqs.max()) >= 2 ** 31:
            raise ValueError("item ids do not fit int32")
        self.seqs = seqs.to(torch.int32).contiguous().to(device)
        self.labels = labels.view(-1).to(torch.int64).contiguous().to(device)
        return self

    @classmethod
    def from_reference_dataloader(cls, dl, mode: str, device="cuda", batch_size: Optional[int] = None):
        """dl: the reference's LRUDataloader (dataloader/lru.py:12-61); mode 'val' or 'test'."""
        if mode not in ("val", "test"):
            raise ValueError("mode must be 'val' or 'test'")
        bs = batch_size or (dl.args.val_batch_size if mode == "val" else dl.args.test_batch_size)
        if mode == "val":
            return cls(dl.train, dl.val, dl.max_len, bs, device=device)
        return cls(dl.train, dl.test, dl.max_len, bs, u2val=dl.val, device=device)

    def __len__(self) -> int:                                      # number of batches, like a DataLoader
        return (len(self.users) + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        for lo in range(0, len(self.users), self.batch_size):
            hi = min(lo + self.batch_size, len(self.users))
            yield self.seqs[lo:hi], self.labels[lo:hi].view(-1, 1)

    def id_bytes(self) -> Dict[str, int]:
        """Bytes of the id matrix here and as the reference's int64 LongTensors."""
        return {"int32_device": self.seqs.numel() * 4, "int64_reference": self.seqs.numel() * 8}
