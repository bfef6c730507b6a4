"""Host-side throughput of the input feed: SFPK shard batches vs the reference pipe restated (torch.load(mmap) per sample
+ normalize + pad/concatenate), same files, page cache warm.  CPU only.  python tools/shard_bench.py [n_samples] [seq] [H]"""
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import offline_feed_oracle as FO                     # noqa: E402  (baseline arm only)
from specforge_b200.shards import Eagle3ShardLoader, list_feature_files, pack_offline_dir   # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
    B = 8
    with tempfile.TemporaryDirectory() as d:
        g = torch.Generator().manual_seed(0)
        for i in range(n):
            L = S + int(torch.randint(0, 200, (1,), generator=g))
            rec = {"input_ids": torch.randint(0, 150000, (L,), generator=g), "loss_mask": torch.ones(L, dtype=torch.int64),
                   "hidden_state": torch.randn(1, L, H, generator=g).bfloat16(), "aux_hidden_state": torch.randn(1, L, 3 * H, generator=g).bfloat16()}
            os.makedirs(os.path.join(d, "feat", "rows_0-2000"), exist_ok=True)
            torch.save(rec, os.path.join(d, "feat", "rows_0-2000", f"data_{i}.ckpt"))
        t0 = time.time()
        (shard,) = pack_offline_dir(os.path.join(d, "feat"), os.path.join(d, "a.sfpk"))
        t_pack = time.time() - t0
        files = list_feature_files(os.path.join(d, "feat"))
        per_batch = B * S * 4 * H * 2

        def ref_epoch():
            for i in range(0, n - B + 1, B):
                raws = [torch.load(p, weights_only=False, mmap=True) for p in files[i:i + B]]
                batch = FO.collate_with_padding([FO.normalize_offline_sample(r, S) for r in raws])
                _ = {k: v.pin_memory() if torch.cuda.is_available() else v for k, v in batch.items()}

        loaders = {t: Eagle3ShardLoader([shard], batch_size=B, max_len=S, threads=t, buffers=2) for t in (1, 8, 16)}

        def ours_epoch(threads):          # the loader (and its ring of destination buffers) lives across epochs, as in training
            for _ in loaders[threads]:
                pass

        for name, fn in (("reference pipe (torch.load mmap + normalize + collate)", ref_epoch),
                         ("SFPK read_batch, 1 thread", lambda: ours_epoch(1)), ("SFPK read_batch, 8 threads", lambda: ours_epoch(8)),
                         ("SFPK read_batch, 16 threads", lambda: ours_epoch(16))):
            fn(); fn()
            t0 = time.time()
            reps = 3
            for _ in range(reps):
                fn()
            dt = (time.time() - t0) / reps
            nb = (n // B)
            print(f"{name:58s} {nb * per_batch / dt / 1e9:7.2f} GB/s  {nb * B / dt:7.1f} samples/s")
        print(f"pack: {n} samples in {t_pack:.2f}s")


if __name__ == "__main__":
    main()
