"""Packed shards of offline feature records + the batch loader that feeds the step.

The reference keeps one `torch.save` file per sample (`scripts/prepare_hidden_states.py:446-480`, directory scheme
`:578-595`), lists them into `SampleRef`s (`runtime/data_plane/offline_reader.py:73-140`), loads each with
`torch.load(mmap=True)` (`feature_store.py:235-240`), truncates and renames per sample
(`algorithms/eagle3/data.py:10-27`) and pads + concatenates on the host (`data/utils.py:106-200`).  Here the same
records live back to back in an `SFPK` shard (layout: `csrc/sf_shard.cpp`) and `sf_shard_read_batch` preads a whole
batch straight into pinned batch-major tensors — truncation, padding and collation included — so the only work left
in Python is the two [B, S] mask fix-ups.  The records are bit-identical to the reference files' tensors and the
batches equal `DataCollatorWithPadding()([normalize_offline_sample(r, max_len) for r in records])`
(tests/test_shards.py checks both, against the reference's own functions when it is importable).
"""
from __future__ import annotations

import ctypes
import os
import struct
import zlib
from typing import Dict, Iterable, Iterator, List, Mapping, Optional, Sequence, Tuple

import torch

from .feed import wait_until_copied

from ._lib import check, lib
from .contracts import TrainBatch

MAGIC = b"SFPK"
VERSION = 1
HEADER_BYTES, FEATURE_BYTES, INDEX_BYTES = 128, 64, 16
_DT = {torch.uint8: 0, torch.int32: 1, torch.int64: 2, torch.bfloat16: 3, torch.float16: 4, torch.float32: 5, torch.bool: 6}
_DT_INV = {v: k for k, v in _DT.items()}
# raw keys of a SpecForge offline EAGLE3 feature file (offline_reader.py:45)
EAGLE3_KEYS = ("input_ids", "loss_mask", "hidden_state", "aux_hidden_state")


def _align(v: int, a: int) -> int:
    return (v + a - 1) // a * a


def _as_rows(name: str, t: torch.Tensor) -> torch.Tensor:
    """[L], [1, L], [L, W] or [1, L, W] -> contiguous [L, W] (the reference stores hidden states with a leading 1)."""
    if t.dim() >= 2 and t.shape[0] == 1 and (t.dim() == 3 or name in ("input_ids", "loss_mask")):
        t = t.squeeze(0)
    if t.dim() == 1:
        t = t.unsqueeze(1)
    if t.dim() != 2:
        raise ValueError(f"feature {name!r}: cannot interpret shape {tuple(t.shape)} as [tokens, width]")
    return t.contiguous()


class ShardWriter:
    """Append records (dicts of tensors keyed by the raw feature names) to one SFPK file."""

    def __init__(self, path: str, features: Sequence[Tuple[str, torch.dtype, int]]):
        if not features:
            raise ValueError("at least one feature")
        for name, dt, width in features:
            if dt not in _DT or width <= 0 or len(name.encode()) > 39:
                raise ValueError(f"bad feature spec {(name, dt, width)}")
        self.path = path
        self.features = [(n, dt, int(w)) for n, dt, w in features]
        self._index: List[Tuple[int, int, int]] = []
        self._f = open(path, "wb")
        self._data_off = _align(HEADER_BYTES + FEATURE_BYTES * len(self.features), 4096)
        self._f.write(b"\0" * self._data_off)
        self._pos = self._data_off

    @classmethod
    def for_record(cls, path: str, record: Mapping[str, torch.Tensor], keys: Sequence[str] = EAGLE3_KEYS) -> "ShardWriter":
        feats = []
        for k in keys:
            rows = _as_rows(k, record[k])
            feats.append((k, rows.dtype, rows.shape[1]))
        return cls(path, feats)

    def add(self, record: Mapping[str, torch.Tensor]) -> int:
        blocks, tokens = [], None
        for name, dt, width in self.features:
            if name not in record:
                raise KeyError(f"record is missing feature {name!r}")
            rows = _as_rows(name, record[name])
            if rows.dtype != dt or rows.shape[1] != width:
                raise ValueError(f"feature {name!r}: expected {dt} x{width}, got {rows.dtype} x{rows.shape[1]}")
            if (torch.is_floating_point(rows)) and bool(torch.isnan(rows).any()):
                raise ValueError(f"NaN in feature {name!r}")       # the reference skips such records (prepare_hidden_states.py:462-470)
            if tokens is None:
                tokens = rows.shape[0]
            elif rows.shape[0] != tokens:
                raise ValueError(f"feature {name!r} has {rows.shape[0]} tokens, others {tokens}")
            blocks.append(rows.cpu())
        off, crc, n = self._pos, 0, 0
        for rows in blocks:    # every block ends on a 64-byte boundary, so block_offset() in the reader is closed-form
            if rows.numel():
                mv = memoryview(rows.view(torch.uint8).numpy()).cast("B")
                self._f.write(mv)
                crc = zlib.crc32(mv, crc)
                n += len(mv)
            tail = b"\0" * (_align(n, 64) - n)
            self._f.write(tail)
            crc = zlib.crc32(tail, crc)
            n += len(tail)
        self._pos += n
        pad = _align(self._pos, 4096) - self._pos
        self._f.write(b"\0" * pad)
        self._pos += pad
        self._index.append((off, int(tokens), crc & 0xFFFFFFFF))
        return len(self._index) - 1

    def close(self) -> None:
        if self._f is None:
            return
        index_off = self._pos
        for off, tokens, crc in self._index:
            self._f.write(struct.pack("<QII", off, tokens, crc))
        file_bytes = index_off + INDEX_BYTES * len(self._index)
        hdr = struct.pack("<4sIIIQQQQQ", MAGIC, VERSION, len(self.features), 0, len(self._index), HEADER_BYTES, index_off,
                          self._data_off, file_bytes)
        self._f.seek(0)
        self._f.write(hdr + b"\0" * (HEADER_BYTES - len(hdr)))
        for name, dt, width in self.features:
            self._f.write(struct.pack("<40sIIQQ", name.encode(), _DT[dt], torch.empty(0, dtype=dt).element_size(), width, 0))
        self._f.close()
        self._f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def list_feature_files(path: str) -> List[str]:
    """Same deterministic order as the reference (offline_reader.py:73-84)."""
    if os.path.isfile(path):
        return [os.path.abspath(path)]
    files = []
    for root, _dirs, names in os.walk(path):
        files += [os.path.abspath(os.path.join(root, n)) for n in names if n.endswith((".ckpt", ".ckpt.gz"))]
    files.sort()
    return files


def pack_offline_dir(src: str, out: str, keys: Sequence[str] = EAGLE3_KEYS, records_per_shard: int = 0) -> List[str]:
    """Repack a directory of reference feature files into SFPK shard(s).  Record i of the concatenated shards is file i
    of the reference's sorted listing, so `SampleRef.sample_id = f"{run_id}:{i:08d}"` keeps its meaning."""
    import gzip
    import io
    files = list_feature_files(src)
    if not files:
        raise FileNotFoundError(f"no .ckpt/.ckpt.gz feature files under {src}")
    outs, writer, n_in = [], None, 0
    for path in files:
        if path.endswith(".gz"):
            with gzip.open(path, "rb") as f:
                raw = torch.load(io.BytesIO(f.read()), weights_only=False)
        else:
            raw = torch.load(path, weights_only=False, mmap=True)
        if writer is None or (records_per_shard and n_in == records_per_shard):
            if writer is not None:
                writer.close()
            name = out if not records_per_shard else f"{out}.{len(outs):05d}"
            writer, n_in = ShardWriter.for_record(name, raw, keys), 0
            outs.append(name)
        writer.add(raw)
        n_in += 1
    writer.close()
    return outs


class ShardReader:
    def __init__(self, path: str):
        L = lib()
        L.sf_shard_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
        L.sf_shard_close.argtypes = [ctypes.c_void_p]
        L.sf_shard_close.restype = None
        L.sf_shard_num_records.argtypes = [ctypes.c_void_p]
        L.sf_shard_num_records.restype = ctypes.c_int64
        L.sf_shard_num_features.argtypes = [ctypes.c_void_p]
        L.sf_shard_feature_info.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int),
                                            ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int64)]
        L.sf_shard_record_tokens.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        L.sf_shard_record_tokens.restype = ctypes.c_int64
        L.sf_shard_verify_record.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        L.sf_shard_read_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.c_int, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
        self._L = L
        self.path = path
        h = ctypes.c_void_p()
        check(L.sf_shard_open(path.encode(), ctypes.byref(h)), "sf_shard_open")
        self._h = h
        self.num_records = int(L.sf_shard_num_records(h))
        self.features: List[Tuple[str, torch.dtype, int]] = []
        for f in range(L.sf_shard_num_features(h)):
            name = ctypes.create_string_buffer(40)
            dt, eb, w = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
            check(L.sf_shard_feature_info(h, f, name, ctypes.byref(dt), ctypes.byref(eb), ctypes.byref(w)), "sf_shard_feature_info")
            self.features.append((name.value.decode(), _DT_INV[dt.value], int(w.value)))

    def close(self) -> None:
        if self._h is not None:
            self._L.sf_shard_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return self.num_records

    def tokens(self, record: int) -> int:
        n = int(self._L.sf_shard_record_tokens(self._h, record))
        if n < 0:
            check(n, "sf_shard_record_tokens")
        return n

    def verify(self, record: int) -> bool:
        rc = self._L.sf_shard_verify_record(self._h, record)
        if rc < 0:
            check(rc, "sf_shard_verify_record")
        return rc == 0

    def read_batch(self, records: Sequence[int], max_tokens: int, pad_tokens: Optional[int] = None,
                   want: Optional[Iterable[str]] = None, pin: bool = True, threads: int = 8,
                   out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """Raw features of `records` as [B, pad_tokens(, width)] host tensors (pinned when CUDA is there); each record is
        cut to its first `max_tokens` tokens and zero-padded.  pad_tokens=None pads to the longest record of the batch
        (what the reference collator does)."""
        n = len(records)
        if pad_tokens is None:
            pad_tokens = max(min(self.tokens(r), max_tokens) for r in records)
        want = set(want) if want is not None else None
        pin = pin and torch.cuda.is_available()
        tensors: Dict[str, torch.Tensor] = {}
        ptrs = (ctypes.c_void_p * len(self.features))()
        for f, (name, dt, width) in enumerate(self.features):
            if want is not None and name not in want:
                ptrs[f] = None
                continue
            shape = (n, pad_tokens) if width == 1 else (n, pad_tokens, width)
            t = out.get(name) if out else None
            if t is None or tuple(t.shape) != shape or t.dtype != dt:
                t = torch.empty(shape, dtype=dt, pin_memory=pin)
            tensors[name] = t
            ptrs[f] = t.data_ptr() if t.numel() else None
        rec = (ctypes.c_int64 * n)(*[int(r) for r in records])
        check(self._L.sf_shard_read_batch(self._h, rec, n, int(max_tokens), int(pad_tokens), ptrs, int(threads)), "sf_shard_read_batch")
        return tensors


class Eagle3ShardLoader:
    RAW_KEYS = EAGLE3_KEYS
    STRATEGY = "eagle3"
    """Iterates TrainBatches for the EAGLE3 offline strategy from SFPK shard(s).

    Batch contents follow the reference exactly: per sample `normalize_offline_sample` (first max_len tokens; target <-
    hidden_state, hidden_state <- aux_hidden_state; last kept loss-mask position zeroed; attention mask of ones,
    algorithms/eagle3/data.py:10-27), then `DataCollatorWithPadding` (zero padding to the longest sample of the batch,
    data/utils.py:106-200).  `pad_to` fixes the padded length instead (e.g. the engine's bound seq_len)."""

    def __init__(self, shards: Sequence[str], batch_size: int, max_len: int, *, run_id: str = "offline", shuffle: bool = False,
                 seed: int = 0, drop_last: bool = True, pad_to: Optional[int] = None, rank: int = 0, world: int = 1,
                 threads: int = 8, pin: bool = True, buffers: int = 6, prefetch: int = 2):
        self.readers = [ShardReader(p) for p in ([shards] if isinstance(shards, str) else shards)]
        self.batch_size, self.max_len, self.run_id = batch_size, max_len, run_id
        self.shuffle, self.seed, self.drop_last, self.pad_to = shuffle, seed, drop_last, pad_to
        self.rank, self.world, self.threads, self.pin = rank, world, threads, pin
        self.prefetch = max(0, prefetch)
        buffers = max(buffers, self.prefetch + 4)     # queue + one in production + DevicePrefetcher depth 2 + the batch in use
        # Ring of reusable (pinned) destination buffer sets: no page faults / cudaHostAlloc per batch.  A yielded batch's host
        # tensors stay valid until `buffers - 1` further batches have been produced (DevicePrefetcher depth 2 needs >= 3).
        self._ring: List[Dict[str, torch.Tensor]] = [dict() for _ in range(max(1, buffers))]
        self._ring_pos = 0
        self.epoch = 0
        self._where: List[Tuple[int, int]] = [(i, r) for i, rd in enumerate(self.readers) for r in range(len(rd))]
        for rd in self.readers:
            names = {n for n, _, _ in rd.features}
            if not set(self.RAW_KEYS) <= names:
                raise KeyError(f"{rd.path}: shard lacks features {sorted(set(self.RAW_KEYS) - names)}")

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def _per_rank(self) -> int:
        """Samples every rank iterates per epoch: ceil(N / world), the DistributedSampler rule the reference's partition follows
        (launch.py:174-239: strided slices of a seeded permutation, wrap-padded) — identical on all ranks, so no rank can run an
        extra optimizer step and block the others in the gradient all-reduce."""
        return (len(self._where) + self.world - 1) // self.world

    def __len__(self) -> int:
        n = self._per_rank()
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _order(self) -> List[int]:
        order = list(range(len(self._where)))
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            order = torch.randperm(len(order), generator=g).tolist()
        total = self._per_rank() * self.world
        if order and len(order) < total:             # wrap-padding: repeat from the start until every rank has the same count
            order = (order * ((total + len(order) - 1) // len(order)))[:total]
        mine = order[self.rank::self.world]          # independent samples: ranks take strided slices
        assert len(mine) == self._per_rank()
        return mine

    def batch(self, sample_indices: Sequence[int]) -> TrainBatch:
        by_reader: Dict[int, List[int]] = {}
        for pos, gi in enumerate(sample_indices):
            by_reader.setdefault(self._where[gi][0], []).append(pos)
        lens = [min(self.readers[self._where[gi][0]].tokens(self._where[gi][1]), self.max_len) for gi in sample_indices]
        S = self.pad_to if self.pad_to is not None else max(lens)
        B = len(sample_indices)
        if len(by_reader) == 1:
            (ri, _), = by_reader.items()
            slot = self._ring[self._ring_pos]
            self._ring_pos = (self._ring_pos + 1) % len(self._ring)
            for t in slot.values():                  # a consumer may still have an asynchronous H2D copy out of this slot in flight
                wait_until_copied(t)
            raw = self.readers[ri].read_batch([self._where[gi][1] for gi in sample_indices], self.max_len, S, self.RAW_KEYS,
                                              self.pin, self.threads, out=slot)
            slot.update(raw)
        else:                                        # a batch straddling shards: gather per shard, then place the rows
            raw = {}
            for ri, poss in by_reader.items():
                part = self.readers[ri].read_batch([self._where[sample_indices[p]][1] for p in poss], self.max_len, S, self.RAW_KEYS,
                                                   False, self.threads)
                for k, v in part.items():
                    if k not in raw:
                        raw[k] = torch.empty((B,) + tuple(v.shape[1:]), dtype=v.dtype, pin_memory=self.pin and torch.cuda.is_available())
                    raw[k][poss] = v
        ids = [f"{self.run_id}:{gi:08d}" for gi in sample_indices]
        tensors, metadata = self._finish(raw, torch.tensor(lens, dtype=torch.int64), S)
        return TrainBatch(sample_ids=ids, strategy=self.STRATEGY, tensors=tensors, metadata=metadata)

    def _finish(self, raw: Dict[str, torch.Tensor], lens_t: torch.Tensor, S: int):
        """Per-sample normalisation left after the gather (algorithms/eagle3/data.py:10-27): rename, attention mask of ones
        over each sample's tokens, last kept loss-mask position zeroed."""
        pos = torch.arange(S, dtype=torch.int64).unsqueeze(0)
        attention_mask = (pos < lens_t.unsqueeze(1)).to(torch.int64)
        loss_mask = raw["loss_mask"]
        live = lens_t > 0
        if bool(live.any()):                         # loss_mask[0, -1] = 0 on the truncated sample (data.py:19-21)
            loss_mask[live.nonzero().squeeze(1), (lens_t[live] - 1)] = 0
        tensors = {"input_ids": raw["input_ids"], "attention_mask": attention_mask, "loss_mask": loss_mask,
                   "hidden_state": raw["aux_hidden_state"], "target": raw["hidden_state"]}
        return tensors, {"target_repr": "hidden_state"}

    def _batches(self) -> Iterator[TrainBatch]:
        order = self._order()
        for i in range(0, len(order), self.batch_size):
            chunk = order[i:i + self.batch_size]
            if len(chunk) < self.batch_size and self.drop_last:
                return
            yield self.batch(chunk)

    def __iter__(self) -> Iterator[TrainBatch]:
        """With prefetch > 0 a background thread gathers the next batches (the C call releases the GIL) while the caller
        launches / waits on the current step, so disk + memcpy time hides behind the GPU."""
        if self.prefetch == 0:
            yield from self._batches()
            return
        import queue
        import threading
        q: "queue.Queue" = queue.Queue(maxsize=self.prefetch)
        stop = threading.Event()

        def produce():
            try:
                for b in self._batches():
                    while not stop.is_set():
                        try:
                            q.put(b, timeout=0.1)
                            break
                        except queue.Full:
                            continue
                    if stop.is_set():
                        return
                q.put(None)
            except BaseException as exc:           # surface reader errors in the consumer
                q.put(exc)

        t = threading.Thread(target=produce, name="sfpk-prefetch", daemon=True)
        t.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            stop.set()
            t.join(timeout=5)


DFLASH_KEYS = ("input_ids", "loss_mask", "hidden_states")


class DFlashShardLoader(Eagle3ShardLoader):
    """TrainBatches for the DFlash-family offline strategy: `normalize_offline_sample` (first max_len tokens of input_ids /
    loss_mask / hidden_states, no renaming, a sample without two consecutive supervised tokens is an error,
    algorithms/common/dflash_family_data.py:37-70) + `pad_and_concatenate_features` (zero padding to the longest sample,
    algorithms/common/collation.py:24-70)."""
    RAW_KEYS = DFLASH_KEYS
    STRATEGY = "dflash"

    def _finish(self, raw: Dict[str, torch.Tensor], lens_t: torch.Tensor, S: int):
        lm = raw["loss_mask"]
        ok = ((lm[:, :-1] > 0) & (lm[:, 1:] > 0)).any(dim=1) if S > 1 else torch.zeros(lm.shape[0], dtype=torch.bool)
        if not bool(ok.all()):
            raise ValueError("offline DFlash-family samples require two consecutive supervised tokens")
        return {"input_ids": raw["input_ids"], "loss_mask": lm, "hidden_states": raw["hidden_states"]}, {}


def _main(argv: Optional[Sequence[str]] = None) -> int:
    """python -m specforge_b200.shards pack <hidden_states_dir> <out.sfpk> [--records-per-shard N]
       python -m specforge_b200.shards info <shard.sfpk> [--verify]"""
    import argparse
    ap = argparse.ArgumentParser(prog="python -m specforge_b200.shards", description="SFPK packed shards of offline feature records")
    sub = ap.add_subparsers(dest="cmd", required=True)
    p = sub.add_parser("pack", help="repack the .ckpt/.ckpt.gz files written by scripts/prepare_hidden_states.py")
    p.add_argument("src")
    p.add_argument("out")
    p.add_argument("--records-per-shard", type=int, default=0)
    q = sub.add_parser("info", help="print the feature table and record lengths of a shard")
    q.add_argument("shard")
    q.add_argument("--verify", action="store_true", help="recompute every record's CRC-32")
    a = ap.parse_args(argv)
    if a.cmd == "pack":
        for name in pack_offline_dir(a.src, a.out, records_per_shard=a.records_per_shard):
            rd = ShardReader(name)
            print(f"{name}: {len(rd)} records, {os.path.getsize(name) / 1e6:.1f} MB")
            rd.close()
        return 0
    rd = ShardReader(a.shard)
    print(f"{a.shard}: {len(rd)} records")
    for name, dt, width in rd.features:
        print(f"  {name:24s} {str(dt):16s} x{width}")
    toks = [rd.tokens(i) for i in range(len(rd))]
    if toks:
        print(f"  tokens/record: min {min(toks)} max {max(toks)} total {sum(toks)}")
    bad = [i for i in range(len(rd)) if a.verify and not rd.verify(i)]
    if a.verify:
        print("  CRC: " + ("all records ok" if not bad else f"MISMATCH in records {bad}"))
    rd.close()
    return 1 if bad else 0


if __name__ == "__main__":
    raise SystemExit(_main())
