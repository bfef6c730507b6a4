"""The drop-in boundary end to end on the GPU, driven by the UNMODIFIED reference package (`baseline/_ref`, vendored onto the GPU
box): the reference's own offline reader / feature store / collator feed the reference's `TrainerCore.train_step`
(training/controller.py:328-363) and `Trainer.fit()` (training/trainer.py:421-431 with the backend and optimizer names swapped by
specforge_b200.train.b200_patches), which drive B200Eagle3TrainStrategy + B200TrainingBackend.  Checks: the loss trace over three
optimizer steps with accumulation against the CPU oracle, the checkpoint the reference's CheckpointManager writes, its export through
the reference's `export_to_sglang`, and a resume.  Tiny dims in the style of tests/test_runtime/_fixtures.py, with head_dim 64 (the
CUDA attention supports 64 / 128; the fixture's 16 is not a production shape)."""
import json
import os
import socket

import pytest
import torch

from _reference import import_reference

pytestmark = pytest.mark.gpu
REF = import_reference()
needs_ref = pytest.mark.skipif(REF is None, reason="reference package not importable (no baseline/_ref, no /root/reference)")

H, V, DV, I, NH, NKV, HD, T = 128, 512, 128, 256, 2, 1, 64, 3
DRAFT_CONFIG = {"architectures": ["LlamaForCausalLMEagle3"], "bos_token_id": 1, "eos_token_id": 2, "hidden_act": "silu", "hidden_size": H,
                "initializer_range": 0.02, "intermediate_size": I, "max_position_embeddings": 512, "model_type": "llama",
                "num_attention_heads": NH, "num_key_value_heads": NKV, "num_hidden_layers": 1, "pad_token_id": 0, "rms_norm_eps": 1e-5,
                "tie_word_embeddings": False, "torch_dtype": "bfloat16", "vocab_size": V, "draft_vocab_size": DV, "head_dim": HD}


@pytest.fixture(scope="module", autouse=True)
def _single_rank_group():
    import torch.distributed as dist
    made = False
    if not dist.is_initialized():
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
        made = True
    yield
    if made:
        dist.destroy_process_group()


def _write_fixtures(work, n=8, seq=40, seed=7):
    """draft.json, a target-head directory, a vocab mapping and n offline feature files (the layout of the reference's
    tests/test_runtime/_fixtures.py writers)."""
    from safetensors.torch import save_file
    os.makedirs(work, exist_ok=True)
    cfg = os.path.join(work, "draft.json")
    with open(cfg, "w") as f:
        json.dump(DRAFT_CONFIG, f)
    tdir = os.path.join(work, "target")
    os.makedirs(tdir, exist_ok=True)
    with open(os.path.join(tdir, "config.json"), "w") as f:
        json.dump({"architectures": ["LlamaForCausalLM"], "model_type": "llama", "hidden_size": H, "vocab_size": V, "num_hidden_layers": 1,
                   "num_attention_heads": 4, "intermediate_size": 128}, f)
    g = torch.Generator().manual_seed(seed)
    head = torch.randn(V, H, generator=g) * 0.3
    save_file({"lm_head.weight": head}, os.path.join(tdir, "model.safetensors"))
    with open(os.path.join(tdir, "model.safetensors.index.json"), "w") as f:
        json.dump({"metadata": {}, "weight_map": {"lm_head.weight": "model.safetensors"}}, f)
    ids = torch.randperm(V, generator=g)[:DV].sort().values
    t2d = torch.zeros(V, dtype=torch.bool)
    t2d[ids] = True
    vmap = os.path.join(work, "vocab_mapping.pt")
    torch.save({"t2d": t2d, "d2t": (ids - torch.arange(DV)).to(torch.int64)}, vmap)
    fdir = os.path.join(work, "features")
    os.makedirs(fdir, exist_ok=True)
    for i in range(n):
        L = seq - (i % 3) * 5                              # ragged lengths: the collator pads each batch to its longest sample
        torch.save({"input_ids": torch.randint(0, V, (L,), generator=g), "loss_mask": torch.ones(L, dtype=torch.long),
                    "hidden_state": torch.randn(1, L, H, generator=g).to(torch.bfloat16),
                    "aux_hidden_state": torch.randn(1, L, 3 * H, generator=g).to(torch.bfloat16)}, os.path.join(fdir, f"{i:04d}.ckpt"))
    return cfg, tdir, vmap, fdir, head


def _build_ours(cfg_path, tdir, vmap, batch, seq):
    from specforge.modeling.target.target_head import TargetHead
    from specforge_b200.draft import B200Eagle3DraftModel
    from specforge_b200.train import B200Eagle3Model
    draft = B200Eagle3DraftModel(DRAFT_CONFIG)
    draft.bind_engine(batch=batch, seq_len=seq, ttt_length=T, device=torch.device("cuda", 0), seed=11)
    draft.load_vocab_mapping(vmap)
    g = torch.Generator().manual_seed(5)
    draft.embed_tokens_weight.data = (torch.randn(V, H, generator=g) * 0.02).to("cuda", torch.bfloat16)
    draft.freeze_embedding()
    model = B200Eagle3Model(draft, length=T, attention_backend="sdpa", lk_loss_type=None, kl_scale=1.0, kl_decay=1.0)
    head = TargetHead.from_pretrained(tdir, lm_head_key="lm_head.weight")
    return draft, model, head


def _oracle_state(draft, head_module):
    from oracle import eagle3_oracle as O
    cfg = O.Eagle3Config(hidden_size=H, intermediate_size=I, num_heads=NH, num_kv_heads=NKV, head_dim=HD, vocab_size=V, draft_vocab_size=DV,
                         rms_norm_eps=1e-5, rope_theta=10000.0, max_position_embeddings=512, ttt_length=T)
    P = {k: v.detach().cpu().clone() for k, v in draft.state_dict().items() if k not in ("t2d", "d2t")}
    return O, cfg, P, head_module.fc.weight.detach().cpu().to(torch.bfloat16), draft.t2d.cpu().clone(), draft.d2t.cpu().clone()


@needs_ref
def test_reference_trainer_core_drives_the_cuda_step(tmp_path):
    """TrainerCore.train_step x 6 micro-batches, accumulation 2 -> 3 optimizer steps; loss trace and grad norms vs the oracle."""
    from specforge.runtime.data_plane import FeatureDataLoader, LocalFeatureStore
    from specforge.training.controller import TrainerCore
    from specforge_b200 import train as T_
    from specforge_b200.backend import B200TrainingBackend
    cfg_path, tdir, vmap, fdir, _ = _write_fixtures(str(tmp_path))
    reg = T_.b200_registry().resolve("eagle3")
    provider = reg.providers.offline_for("text")
    refs = provider.build_reader(fdir, run_id="data", ttt_length=T, max_len=64).read()
    loader = FeatureDataLoader(LocalFeatureStore("data-features"), refs=refs, batch_size=2, collate_fn=provider.build_collator(),
                               per_sample_transform=provider.build_normalizer(64), strategy=reg.name)
    draft, model, head = _build_ours(cfg_path, tdir, vmap, batch=2, seq=64)
    backend = B200TrainingBackend(lr=1e-3, max_grad_norm=0.5, total_steps=20, warmup_ratio=0.1)
    wrapped = backend.prepare_model(model, optimizer_target=model.draft_model)
    strategy = reg.providers.step.build(wrapped, target_head=head, **dict(compact_teacher=False, compact_teacher_chunk_size=None))
    assert backend.strategy is strategy                                  # attached through the draft module
    core = TrainerCore(strategy, backend, accumulation_steps=2)
    O, ocfg, P, head_w, t2d, d2t = _oracle_state(draft, head)
    names = [n for n in O.PARAM_NAMES]
    masters = [P[n].float() for n in names]
    ea, es = [torch.zeros_like(m) for m in masters], [torch.zeros_like(m) for m in masters]
    acc = None
    losses, ref_losses, gnorms, ref_gnorms = [], [], [], []
    batches = list(loader)[:3] + list(loader)[:3]
    from specforge_b200.optimizer import WarmupSchedule
    sched = WarmupSchedule(1e-3, 20, 2, "cosine")
    for i, batch in enumerate(batches):
        res = core.train_step(batch)
        losses.append(float(res.metrics["loss"]))
        if res.grad_norm is not None:
            gnorms.append(float(res.grad_norm))
        cpu_batch = {k: v.clone() for k, v in batch.tensors.items()}
        r, grads = O.train_step(P, ocfg, cpu_batch, head_w, t2d, d2t)
        ref_losses.append(float(r.loss))
        g = [grads[n].float() / 2 for n in names]                        # loss / accumulation_steps
        acc = g if acc is None else [a + b for a, b in zip(acc, g)]
        if i % 2 == 1:
            gb = [a.to(torch.bfloat16) for a in acc]
            ref_gnorms.append(float(torch.sqrt(sum((x.float() ** 2).sum() for x in gb))))
            O.adamw_clip_step([P[n] for n in names], masters, ea, es, gb, step=len(ref_gnorms), lr=sched.lr_at(len(ref_gnorms) - 1))
            acc = None
    torch.cuda.synchronize()
    assert len(gnorms) == 3 and backend.optimizer.scheduler.n == 3 and backend.engine.opt_step == 3
    for a, b in zip(losses, ref_losses):
        assert a == pytest.approx(b, rel=2e-3), (losses, ref_losses)
    for a, b in zip(gnorms, ref_gnorms):
        assert a == pytest.approx(b, rel=2e-2), (gnorms, ref_gnorms)
    # the weights after three steps track the oracle's (AdamW moves each weight by ~lr per step)
    for n in names:
        got, ref = draft.state_dict()[n].float().cpu(), P[n].float()
        assert (got - ref).abs().max().item() <= 4e-3 + 2 ** -7 * ref.abs().max().item(), n


@needs_ref
def test_reference_trainer_fit_checkpoint_export_resume(tmp_path):
    """`build_offline_runtime(...).fit()` — the reference Trainer, controller, loader and CheckpointManager — with the backend /
    optimizer names swapped by b200_patches(); then the reference exporter on the checkpoint it wrote, and a resumed run."""
    from specforge.export.to_sglang import export_to_sglang
    from specforge.launch import build_offline_runtime
    from specforge_b200 import train as T_
    from specforge_b200.backend import B200TrainingBackend
    from specforge_b200.optimizer import B200BF16Optimizer
    cfg_path, tdir, vmap, fdir, _ = _write_fixtures(str(tmp_path), n=8)
    reg = T_.b200_registry().resolve("eagle3")
    out_dir = os.path.join(str(tmp_path), "output")

    def opt_factory(module):
        return B200BF16Optimizer(module, lr=1e-3, max_grad_norm=0.5, warmup_ratio=0.0, total_steps=4)

    def run(max_steps, resume_from=None, seed=11):
        draft, model, head = _build_ours(cfg_path, tdir, vmap, batch=2, seq=64)
        logged = []
        with T_.b200_patches():
            trainer = build_offline_runtime(algorithm=reg, hidden_states_path=fdir, draft_model=model, target_head=head,
                                            optimizer_factory=opt_factory, run_id="b200-offline", output_dir=out_dir, ttt_length=T,
                                            max_len=64, batch_size=2, accumulation_steps=2, max_steps=max_steps, total_steps=4,
                                            save_interval=1, logger=lambda m, s: logged.append((s, m["loss"])), log_interval=1,
                                            resume_from=resume_from)
            assert isinstance(trainer.backend, B200TrainingBackend)          # seam 3: trainer.py:421 built OUR backend
            steps = trainer.fit()
        return steps, logged, draft, trainer

    steps, logged, draft, trainer = run(max_steps=2)
    assert steps == 2 and [s for s, _ in logged] == [1, 2]
    assert all(torch.isfinite(torch.tensor(l)) and 1.0 < l < 20.0 for _, l in logged)
    ckpt = os.path.join(out_dir, "b200-offline-latest")
    assert os.path.exists(ckpt)
    state = torch.load(os.path.join(os.path.realpath(ckpt), "training_state.pt"), weights_only=False) if os.path.exists(
        os.path.join(os.path.realpath(ckpt), "training_state.pt")) else None
    # the reference exporter reads the checkpoint, materialises ITS LlamaForCausalLMEagle3 from our draft_state_dict and writes an
    # SGLang draft directory (export/to_sglang.py:37-90)
    exported = export_to_sglang(ckpt, cfg_path, os.path.join(str(tmp_path), "sglang"), vocab_mapping_path=vmap)
    from safetensors import safe_open
    with safe_open(os.path.join(exported, "model.safetensors"), framework="pt") as f:
        keys = set(f.keys())
        fc = f.get_tensor("fc.weight")
    assert {"fc.weight", "norm.weight", "lm_head.weight", "t2d", "d2t", "midlayer.self_attn.q_proj.weight"} <= keys
    assert not any(k.startswith("draft_model.") or "embed" in k for k in keys)
    assert torch.equal(fc.to(torch.bfloat16).cpu(), draft.state_dict()["fc.weight"].cpu())
    if state is not None:
        opt = state.get("replicated_optimizer_state")
        assert opt is not None and set(opt) == {"optimizer_state_dict", "scheduler_state_dict", "lr_scheduler_type", "max_grad_norm", "fp32_params"}
