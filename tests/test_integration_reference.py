"""Seam 1 (draft registry) against the REAL reference package, when it is present (/root/reference exists only in the
build container; on the GPU box this module is skipped).  CPU only: checks that our draft class can be bound into the
reference's DRAFT_REGISTRY, resolved by AutoDraftModelConfig / AutoDraftModel from the reference's own draft JSONs, and
that its state-dict contract (names + shapes) equals the reference class's, incl. EAGLE3.1 fc_norm."""
import os
import sys

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "specforge")), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import specforge.modeling.draft.llama3_eagle as le
    from specforge.modeling.auto import AutoDraftModel, AutoDraftModelConfig
    from specforge.modeling.draft.registry import DRAFT_REGISTRY
    return le, AutoDraftModel, AutoDraftModelConfig, DRAFT_REGISTRY


@pytest.mark.parametrize("cfg_name", ["qwen2.5-0.5b-eagle3.json", "qwen3-30B-A3B-eagle3.1.json"])
def test_registry_binding_and_state_dict_contract(ref, cfg_name, monkeypatch):
    le, AutoDraftModel, AutoDraftModelConfig, DRAFT_REGISTRY = ref
    from specforge_b200.draft import B200Eagle3DraftModel
    ref_cls = le.LlamaForCausalLMEagle3
    path = os.path.join(REF, "configs", cfg_name)
    config = AutoDraftModelConfig.from_file(path)
    # shrink the tables so the reference module instantiates quickly on CPU (shapes scale with these fields only)
    config.vocab_size, config.draft_vocab_size = 2048, 512
    ref_model = ref_cls(config, attention_backend="sdpa")
    ref_sd = {k: tuple(v.shape) for k, v in ref_model.state_dict().items()}
    monkeypatch.setitem(DRAFT_REGISTRY, "LlamaForCausalLMEagle3", B200Eagle3DraftModel)   # INTEGRATION.md seam 1
    config2 = AutoDraftModelConfig.from_file(path)                    # resolves config_class through OUR class
    config2.vocab_size, config2.draft_vocab_size = 2048, 512
    ours = AutoDraftModel.from_config(config2, attention_backend="b200")
    assert isinstance(ours, B200Eagle3DraftModel)
    assert ours.state_dict_spec() == ref_sd
    assert ours.dims.fc_norm == bool(getattr(config, "fc_norm", False))


def test_shard_batches_equal_the_reference_input_pipe(ref, tmp_path):
    """Input-feed row: SFPK shard batches == the reference's OfflineManifestReader listing + load_feature_file +
    normalize_offline_sample + DataCollatorWithPadding on the same files; also pins oracle/offline_feed_oracle.py."""
    import torch
    from specforge.algorithms.eagle3.data import normalize_offline_sample
    from specforge.data.utils import DataCollatorWithPadding
    from specforge.runtime.data_plane.feature_store import load_feature_file
    from specforge.runtime.data_plane.offline_reader import OfflineManifestReader
    from oracle import offline_feed_oracle as FO
    from specforge_b200.shards import Eagle3ShardLoader, pack_offline_dir
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_shards import _write_reference_files
    _write_reference_files(str(tmp_path / "feat"), [37, 5, 64, 1, 90, 12, 33, 48], seed=9)
    reader = OfflineManifestReader(str(tmp_path / "feat"), run_id="offline", max_len=48)
    refs = reader.read()
    assert len(refs) == 8
    (shard,) = pack_offline_dir(str(tmp_path / "feat"), str(tmp_path / "a.sfpk"))
    collate = DataCollatorWithPadding()
    for bi, batch in enumerate(Eagle3ShardLoader([shard], batch_size=4, max_len=48)):
        rs = refs[4 * bi:4 * bi + 4]
        assert batch.sample_ids == [r.sample_id for r in rs]
        raws = [load_feature_file(r.feature_store_uri[len("file://"):]) for r in rs]
        want = collate([normalize_offline_sample(raw, 48) for raw in raws])
        mine = FO.collate_with_padding([FO.normalize_offline_sample(raw, 48) for raw in raws])
        for k in ("input_ids", "attention_mask", "loss_mask", "hidden_state", "target"):
            w, g, o = want[k], batch.tensors[k], mine[k]
            assert g.dtype == w.dtype and g.shape == w.shape and o.shape == w.shape, k
            same = lambda a, b: torch.equal(a.view(torch.int16), b.view(torch.int16)) if a.dtype == torch.bfloat16 else torch.equal(a, b)
            assert same(g, w) and same(o, w), k


def test_dflash_registry_binding_and_state_dict_contract(ref, monkeypatch):
    """Seam 1 for the DFlash draft: our class resolves from the reference's own draft JSON through its registry and promises
    exactly the reference DFlashDraftModel's parameter names and shapes."""
    le, AutoDraftModel, AutoDraftModelConfig, DRAFT_REGISTRY = ref
    from specforge.modeling.draft.dflash import DFlashDraftModel
    from specforge_b200.dflash import B200DFlashDraftModel
    path = os.path.join(REF, "configs", "qwen3-8b-dflash.json")
    config = AutoDraftModelConfig.from_file(path)
    config.vocab_size, config.hidden_size, config.intermediate_size = 1024, 256, 512      # shrink: shapes scale with these only
    config.num_attention_heads, config.num_key_value_heads, config.head_dim = 8, 2, 32
    ref_sd = {k: tuple(v.shape) for k, v in DFlashDraftModel(config).state_dict().items()}
    monkeypatch.setitem(DRAFT_REGISTRY, "DFlashDraftModel", B200DFlashDraftModel)
    ours = AutoDraftModel.from_config(config)
    assert isinstance(ours, B200DFlashDraftModel)
    assert ours.state_dict_spec() == ref_sd
    assert ours.block_size == 16 and ours.mask_token_id == 151669 and list(ours.target_layer_ids) == [1, 9, 17, 25, 33]
    assert ours.dims.num_target_feats == 5 and ours.dims.num_layers == 5


def test_dflash_shard_batches_equal_the_reference_input_pipe(ref, tmp_path):
    """DFlash-family offline pipe: SFPK batches == the reference's normalize_offline_sample + build_collator() on the same
    files (algorithms/common/dflash_family_data.py); also pins the oracle restatement."""
    import torch
    from specforge.algorithms.common.dflash_family_data import build_collator, normalize_offline_sample
    from specforge.runtime.data_plane.feature_store import load_feature_file
    from oracle import offline_feed_oracle as FO
    from specforge_b200.shards import DFLASH_KEYS, DFlashShardLoader, list_feature_files, pack_offline_dir
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_shards import _write_dflash_files
    _write_dflash_files(str(tmp_path / "feat"), [30, 7, 52, 19, 40, 33], seed=6)
    files = list_feature_files(str(tmp_path / "feat"))
    (shard,) = pack_offline_dir(str(tmp_path / "feat"), str(tmp_path / "d.sfpk"), keys=DFLASH_KEYS)
    collate = build_collator()
    same = lambda a, b: torch.equal(a.view(torch.int16), b.view(torch.int16)) if a.dtype == torch.bfloat16 else torch.equal(a, b)
    for bi, batch in enumerate(DFlashShardLoader([shard], batch_size=3, max_len=36)):
        raws = [load_feature_file(p) for p in files[3 * bi:3 * bi + 3]]
        want = collate([normalize_offline_sample(r, 36) for r in raws])
        mine = FO.pad_and_concatenate([FO.normalize_offline_dflash_sample(r, 36) for r in raws])
        for k in ("input_ids", "loss_mask", "hidden_states"):
            assert batch.tensors[k].shape == want[k].shape and batch.tensors[k].dtype == want[k].dtype, k
            assert same(batch.tensors[k], want[k]) and same(mine[k], want[k]), k


def test_checkpoint_keys_satisfy_the_reference_sglang_exporter(ref):
    """What `backend.state_dict()["model"]` holds (reference-named draft weights, embedding dropped by
    checkpoint_state_filter) passes the exporter's serving-key validation (export/to_sglang.py:37-53)."""
    import torch
    from specforge.export.to_sglang import WEIGHT_MAPS, _serving_state
    from specforge_b200.draft import B200Eagle3DraftModel
    cfg = dict(hidden_size=256, intermediate_size=512, num_attention_heads=4, num_key_value_heads=2, head_dim=64, vocab_size=1024,
               draft_vocab_size=256, rms_norm_eps=1e-6, rope_theta=10000.0, max_position_embeddings=512)
    spec = B200Eagle3DraftModel(cfg).state_dict_spec()
    trainer_side = {"draft_model." + k: torch.empty(0) for k in spec}                       # what the trainable module's state_dict carries
    kept = {k.replace("draft_model.", ""): v for k, v in trainer_side.items() if "embed" not in k.lower()}   # checkpoint_state_filter
    out = _serving_state(kept, WEIGHT_MAPS["LlamaForCausalLMEagle3"])
    assert {"fc.weight", "norm.weight", "lm_head.weight", "t2d", "d2t"} <= set(out) and "embed_tokens.weight" not in out
    with pytest.raises(ValueError):
        _serving_state(trainer_side, {})                                                       # unfiltered keys are refused
