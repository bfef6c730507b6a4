"""SFPK packed shards (CPU): records bit-identical to the reference-format files they were packed from, batches equal to
normalize_offline_sample + DataCollatorWithPadding (oracle restatement; the real reference functions are used in
test_integration_reference.py), truncation / padding / empty / multi-shard / corruption edge cases."""
import os

import pytest
import torch

from oracle import offline_feed_oracle as FO


def _write_reference_files(root, lengths, H=16, seed=0, group=4):
    """The reference's layout: rows_{g}-{g+group}/data_{i}.ckpt, one torch.save dict per sample
    (prepare_hidden_states.py:446-480, :578-595)."""
    g = torch.Generator().manual_seed(seed)
    recs = []
    for i, L in enumerate(lengths):
        rec = {"input_ids": torch.randint(0, 1000, (L,), generator=g),
               "loss_mask": torch.randint(0, 2, (L,), generator=g),
               "hidden_state": torch.randn(1, L, H, generator=g).bfloat16(),
               "aux_hidden_state": torch.randn(1, L, 3 * H, generator=g).bfloat16()}
        gi = (i // group) * group
        d = os.path.join(root, f"rows_{gi}-{gi + group}")
        os.makedirs(d, exist_ok=True)
        torch.save(rec, os.path.join(d, f"data_{i}.ckpt"))
        recs.append(rec)
    return recs


def _ref_order(root):
    from specforge_b200.shards import list_feature_files
    return [int(os.path.basename(p)[5:-5]) for p in list_feature_files(root)]


def test_pack_and_read_records_bit_exact(tmp_path):
    from specforge_b200.shards import ShardReader, pack_offline_dir
    lengths = [37, 5, 64, 1, 90, 12, 33]
    recs = _write_reference_files(str(tmp_path / "feat"), lengths)
    order = _ref_order(str(tmp_path / "feat"))          # lexicographic, like the reference's sorted listing
    (shard,) = pack_offline_dir(str(tmp_path / "feat"), str(tmp_path / "a.sfpk"))
    rd = ShardReader(shard)
    assert len(rd) == len(lengths)
    assert [n for n, _, _ in rd.features] == ["input_ids", "loss_mask", "hidden_state", "aux_hidden_state"]
    for r, src in enumerate(order):
        assert rd.tokens(r) == lengths[src] and rd.verify(r)
        got = rd.read_batch([r], max_tokens=10 ** 6)
        assert torch.equal(got["input_ids"][0], recs[src]["input_ids"])
        assert torch.equal(got["loss_mask"][0], recs[src]["loss_mask"])
        assert torch.equal(got["hidden_state"][0].view(torch.int16), recs[src]["hidden_state"][0].view(torch.int16))
        assert torch.equal(got["aux_hidden_state"][0].view(torch.int16), recs[src]["aux_hidden_state"][0].view(torch.int16))


@pytest.mark.parametrize("max_len,pad_to", [(48, None), (48, 64), (1000, None)])
def test_batches_equal_reference_normalise_and_collate(tmp_path, max_len, pad_to):
    from specforge_b200.shards import Eagle3ShardLoader, pack_offline_dir
    lengths = [37, 5, 64, 1, 90, 12, 33, 48]
    recs = _write_reference_files(str(tmp_path / "feat"), lengths, seed=3)
    order = _ref_order(str(tmp_path / "feat"))
    (shard,) = pack_offline_dir(str(tmp_path / "feat"), str(tmp_path / "a.sfpk"))
    loader = Eagle3ShardLoader([shard], batch_size=4, max_len=max_len, pad_to=pad_to, threads=3)
    assert len(loader) == 2
    for bi, batch in enumerate(loader):
        srcs = order[4 * bi:4 * bi + 4]
        want = FO.collate_with_padding([FO.normalize_offline_sample(recs[s], max_len) for s in srcs])
        S = want["input_ids"].shape[1]
        assert batch.sample_ids == [f"offline:{4 * bi + k:08d}" for k in range(4)]
        assert batch.strategy == "eagle3" and batch.metadata["target_repr"] == "hidden_state"
        for k, w in want.items():
            got = batch.tensors[k]
            if pad_to is not None:
                assert got.shape[1] == pad_to and not got[:, S:].any()      # extra padding is zeros (masks included)
                got = got[:, :S]
            assert got.dtype == w.dtype and got.shape == w.shape, k
            assert torch.equal(got.view(torch.int16) if got.dtype == torch.bfloat16 else got,
                               w.view(torch.int16) if w.dtype == torch.bfloat16 else w), k


def test_multi_shard_shuffle_rank_split_and_empty_record(tmp_path):
    from specforge_b200.shards import Eagle3ShardLoader, ShardReader, ShardWriter, pack_offline_dir
    recs = _write_reference_files(str(tmp_path / "feat"), [20, 30, 0, 25, 31, 7], seed=5)
    shards = pack_offline_dir(str(tmp_path / "feat"), str(tmp_path / "s"), records_per_shard=4)
    assert len(shards) == 2 and len(ShardReader(shards[0])) == 4 and len(ShardReader(shards[1])) == 2
    order = _ref_order(str(tmp_path / "feat"))
    seen = []
    for rank in range(2):
        ld = Eagle3ShardLoader(shards, batch_size=3, max_len=28, shuffle=True, seed=11, rank=rank, world=2, drop_last=False)
        for batch in ld:            # every batch straddles the two shards or not — both paths must agree with the oracle
            idx = [int(s.split(":")[1]) for s in batch.sample_ids]
            seen += idx
            want = FO.collate_with_padding([FO.normalize_offline_sample(recs[order[i]], 28) for i in idx])
            for k, w in want.items():
                g = batch.tensors[k]
                assert torch.equal(g.view(torch.int16) if g.dtype == torch.bfloat16 else g,
                                   w.view(torch.int16) if w.dtype == torch.bfloat16 else w), k
    assert sorted(seen) == list(range(6))          # ranks cover the data set exactly once
    # N not divisible by world: DistributedSampler semantics — every rank iterates ceil(N / world) samples (wrap-padded), so all
    # ranks run the same number of steps (an extra step on one rank would block forever in the gradient all-reduce)
    per_rank = []
    for rank in range(4):
        ld = Eagle3ShardLoader(shards, batch_size=1, max_len=28, shuffle=True, seed=3, rank=rank, world=4, drop_last=True)
        ids = [int(b.sample_ids[0].split(":")[1]) for b in ld]
        assert len(ids) == len(ld) == 2
        per_rank.append(ids)
    assert sorted(set(sum(per_rank, []))) == list(range(6))     # everything is still covered; two samples appear twice
    # writer-side validation
    with pytest.raises(ValueError):
        w = ShardWriter(str(tmp_path / "bad.sfpk"), [("x", torch.float32, 4)])
        w.add({"x": torch.full((3, 4), float("nan"))})


def test_corruption_is_detected(tmp_path):
    from specforge_b200._lib import SfError
    from specforge_b200.shards import ShardReader, pack_offline_dir
    _write_reference_files(str(tmp_path / "feat"), [40, 41], seed=7)
    (shard,) = pack_offline_dir(str(tmp_path / "feat"), str(tmp_path / "a.sfpk"))
    rd = ShardReader(shard)
    assert rd.verify(0) and rd.verify(1)
    rd.close()
    blob = bytearray(open(shard, "rb").read())
    blob[4096 + 100] ^= 0x40                                    # one flipped payload bit in record 0
    open(shard, "wb").write(bytes(blob))
    rd = ShardReader(shard)
    assert not rd.verify(0) and rd.verify(1)
    rd.close()
    open(shard, "wb").write(bytes(blob[:-7]))                   # truncated file
    with pytest.raises(SfError):
        ShardReader(shard)
    open(shard, "wb").write(b"NOPE" + bytes(blob[4:]))          # wrong magic
    with pytest.raises(SfError):
        ShardReader(shard)
    with pytest.raises(SfError):
        ShardReader(str(tmp_path / "missing.sfpk"))


def test_pack_and_info_cli(tmp_path, capsys):
    from specforge_b200.shards import _main
    _write_reference_files(str(tmp_path / "feat"), [9, 20, 3], seed=2)
    assert _main(["pack", str(tmp_path / "feat"), str(tmp_path / "x.sfpk")]) == 0
    assert _main(["info", str(tmp_path / "x.sfpk"), "--verify"]) == 0
    out = capsys.readouterr().out
    assert "3 records" in out and "aux_hidden_state" in out and "all records ok" in out


def _write_dflash_files(root, lengths, W=24, seed=0):
    g = torch.Generator().manual_seed(seed)
    recs = []
    os.makedirs(os.path.join(root, "rows_0-2000"), exist_ok=True)
    for i, L in enumerate(lengths):
        lm = (torch.rand(L, generator=g) > 0.3).long()
        lm[:2] = 1                                                   # two consecutive supervised tokens in every sample
        hs = torch.randn(L, W, generator=g).bfloat16()
        rec = {"input_ids": torch.randint(0, 1000, (L,), generator=g), "loss_mask": lm,
               "hidden_states": hs if i % 2 else hs.unsqueeze(0)}     # both stored layouts ([seq, W] and [1, seq, W])
        torch.save(rec, os.path.join(root, "rows_0-2000", f"data_{i}.ckpt"))
        recs.append(rec)
    return recs


def test_dflash_batches_equal_reference_normalise_and_collate(tmp_path):
    from specforge_b200.shards import DFLASH_KEYS, DFlashShardLoader, pack_offline_dir
    lengths = [30, 7, 52, 19, 40, 33]
    recs = _write_dflash_files(str(tmp_path / "feat"), lengths, seed=4)
    order = _ref_order(str(tmp_path / "feat"))
    (shard,) = pack_offline_dir(str(tmp_path / "feat"), str(tmp_path / "d.sfpk"), keys=DFLASH_KEYS)
    for bi, batch in enumerate(DFlashShardLoader([shard], batch_size=3, max_len=36)):
        srcs = order[3 * bi:3 * bi + 3]
        want = FO.pad_and_concatenate([FO.normalize_offline_dflash_sample(recs[s], 36) for s in srcs])
        assert batch.strategy == "dflash" and set(batch.tensors) == set(want)
        for k, w in want.items():
            g = batch.tensors[k]
            assert g.dtype == w.dtype and g.shape == w.shape, k
            assert torch.equal(g.view(torch.int16) if g.dtype == torch.bfloat16 else g, w.view(torch.int16) if w.dtype == torch.bfloat16 else w), k
    # a sample without two consecutive supervised tokens is refused, like the reference normaliser does
    bad = dict(recs[0]); bad["loss_mask"] = torch.zeros_like(bad["loss_mask"]); bad["loss_mask"][::2] = 1
    os.makedirs(str(tmp_path / "bad" / "rows_0-2000"))
    torch.save(bad, str(tmp_path / "bad" / "rows_0-2000" / "data_0.ckpt"))
    (bshard,) = pack_offline_dir(str(tmp_path / "bad"), str(tmp_path / "b.sfpk"), keys=DFLASH_KEYS)
    with pytest.raises(ValueError):
        next(iter(DFlashShardLoader([bshard], batch_size=1, max_len=36, prefetch=0)))
