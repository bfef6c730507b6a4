"""Host-side pieces of the drop-in boundary against the UNMODIFIED reference package (CPU; skipped where the reference is
not importable): the `specforge train` registry wrapper (seam 4), the Trainer patch points (seam 3), and the optimizer
checkpoint layout — B200BF16Optimizer.state_dict() loads into the reference BF16Optimizer and back, the warm-up scheduler
state matches the reference scheduler's step for step (optimizer.py:170-229, lr_scheduler.py:56-147)."""
import copy

import pytest
import torch

from _cpu_engine import CpuFlatEngine
from _reference import import_reference

REF = import_reference()
needs_ref = pytest.mark.skipif(REF is None, reason="reference package not importable (no baseline/_ref, no /root/reference)")

SHAPES = {   # a tiny LlamaForCausalLMEagle3: H=128, I=256, nh=2, nkv=1, d=64, DV=128
    "fc.weight": (128, 384), "midlayer.self_attn.q_proj.weight": (128, 256), "midlayer.self_attn.k_proj.weight": (64, 256),
    "midlayer.self_attn.v_proj.weight": (64, 256), "midlayer.self_attn.o_proj.weight": (128, 128),
    "midlayer.mlp.gate_proj.weight": (256, 128), "midlayer.mlp.up_proj.weight": (256, 128),
    "midlayer.mlp.down_proj.weight": (128, 256), "midlayer.hidden_norm.weight": (128,), "midlayer.input_layernorm.weight": (128,),
    "midlayer.post_attention_layernorm.weight": (128,), "norm.weight": (128,), "lm_head.weight": (128, 128),
}


@pytest.fixture(scope="module", autouse=True)
def _single_rank_group():
    """The reference's loaders print through print_on_rank0, which needs an initialised process group (gloo, world 1)."""
    import socket
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


class _Holder(torch.nn.Module):
    def __init__(self, engine):
        super().__init__()
        self.engine = engine


def _reference_draft():
    from transformers import LlamaConfig
    from specforge.modeling.draft.llama3_eagle import LlamaForCausalLMEagle3
    hf = LlamaConfig(hidden_size=128, intermediate_size=256, num_attention_heads=2, num_key_value_heads=1, num_hidden_layers=1,
                     vocab_size=512, rms_norm_eps=1e-5, max_position_embeddings=512, hidden_act="silu", tie_word_embeddings=False,
                     pad_token_id=0)
    hf.head_dim, hf.draft_vocab_size = 64, 128
    m = LlamaForCausalLMEagle3(hf, attention_backend="sdpa").to(torch.bfloat16)
    m.freeze_embedding()
    return m


@needs_ref
def test_parameter_order_matches_the_reference_module():
    from specforge_b200.optimizer import REFERENCE_PARAM_ORDER
    m = _reference_draft()
    names = [n for n, p in m.named_parameters() if p.requires_grad]
    assert names == [n for n in REFERENCE_PARAM_ORDER if n in SHAPES]
    assert {n: tuple(p.shape) for n, p in m.named_parameters() if p.requires_grad} == SHAPES


@needs_ref
@pytest.mark.parametrize("kind,total,ratio", [("cosine", 40, 0.25), ("cosine", 30, 0.0), ("constant", 20, 0.2)])
def test_warmup_schedule_tracks_the_reference_scheduler(kind, total, ratio):
    from specforge.optimizer import BF16Optimizer
    from specforge_b200.optimizer import WarmupSchedule
    ref = BF16Optimizer(torch.nn.Linear(2, 2), lr=3e-4, total_steps=total, warmup_ratio=ratio, lr_scheduler=kind)
    ours = WarmupSchedule(3e-4, total, int(ratio * total), kind)
    for n in range(total - 1):
        assert ours.lr == pytest.approx(ref.get_learning_rate(), rel=1e-6, abs=1e-12), n
        rs, os_ = ref.scheduler.state_dict(), ours.state_dict()
        for key in ("warmup_epochs", "finished", "last_epoch", "after_scheduler_type"):
            assert rs[key] == os_[key], (n, key, rs[key], os_[key])
        assert rs["after_scheduler_dict"]["last_epoch"] == os_["after_scheduler_dict"]["last_epoch"], n
        # a reference scheduler restored from OUR dict continues on the same learning rates
        clone = BF16Optimizer(torch.nn.Linear(2, 2), lr=3e-4, total_steps=total, warmup_ratio=ratio, lr_scheduler=kind)
        clone.scheduler.load_state_dict(copy.deepcopy(os_))
        back = WarmupSchedule(3e-4, total, int(ratio * total), kind)
        back.load_state_dict(rs)
        assert back.n == n
        ref.optimizer.step(); ref.scheduler.step(); clone.optimizer.step(); clone.scheduler.step(); ours.step()
        assert clone.get_learning_rate() == pytest.approx(ref.get_learning_rate(), rel=1e-6, abs=1e-12), n


def _same_weights(a, b, what):
    """bf16 weights written back from fp32 masters that differ by fp32 rounding noise (the reference sums per-parameter
    squared norms, the fused step one flat norm): identical except where a master sits on a bf16 rounding boundary."""
    a, b = a.float(), b.float()
    assert (a != b).float().mean().item() < 1e-3, what
    assert (a - b).abs().max().item() <= 2 ** -7 * max(a.abs().max().item(), 1e-6), what


def _random_grads(engine, seed):
    g = torch.Generator().manual_seed(seed)
    engine.grads_f32.copy_(torch.randn(engine.n_params, generator=g) * 1e-2)
    engine.grads_to_bf16()


@needs_ref
def test_optimizer_state_crossloads_with_reference_bf16optimizer():
    """3 steps here -> state_dict -> reference BF16Optimizer.load_state_dict -> both take a 4th step on the same gradient: same
    weights.  And the other way round: the reference's state after 3 steps loads into ours."""
    from specforge.optimizer import BF16Optimizer
    from specforge_b200.optimizer import B200BF16Optimizer
    kw = dict(lr=1e-3, max_grad_norm=0.5, total_steps=50, warmup_ratio=0.1)
    eng = CpuFlatEngine(SHAPES, seed=1)
    ours = B200BF16Optimizer(_Holder(eng), **kw)
    ref_model = _reference_draft()
    with torch.no_grad():
        for n, p in ref_model.named_parameters():
            if p.requires_grad:
                p.copy_(eng.param_view(n))
    ref = BF16Optimizer(ref_model, **kw)
    named = {n: p for n, p in ref_model.named_parameters() if p.requires_grad}
    for step in range(3):
        _random_grads(eng, 10 + step)
        for n, p in named.items():
            p.grad = eng.param_view(n, eng.grads_bf16).clone()
        gn_o, gn_r = ours.step(), ref.step()
        assert float(gn_o) == pytest.approx(float(gn_r), rel=1e-4)
        for n, p in named.items():
            _same_weights(p.detach(), eng.param_view(n), (step, n))
    sd = ours.state_dict()
    assert set(sd) == {"optimizer_state_dict", "scheduler_state_dict", "lr_scheduler_type", "max_grad_norm", "fp32_params"}
    assert set(sd) == set(ref.state_dict()) and len(sd["fp32_params"]) == len(ref.state_dict()["fp32_params"])
    assert set(sd["optimizer_state_dict"]["param_groups"][0]) == set(ref.state_dict()["optimizer_state_dict"]["param_groups"][0])
    # ours -> reference
    ref2_model = _reference_draft()
    ref2 = BF16Optimizer(ref2_model, **kw)
    ref2.load_state_dict(copy.deepcopy(sd))
    with torch.no_grad():
        for n, p in ref2_model.named_parameters():
            if p.requires_grad:
                p.copy_(eng.param_view(n))
    # reference -> ours
    eng3 = CpuFlatEngine(SHAPES, seed=99)
    ours3 = B200BF16Optimizer(_Holder(eng3), **kw)
    eng3.params.copy_(eng.params)
    ours3.load_state_dict(copy.deepcopy(ref.state_dict()))
    assert eng3.opt_step == 3 and ours3.scheduler.n == 3
    _random_grads(eng, 77)
    eng3.grads_bf16.copy_(eng.grads_bf16)
    for model in (ref_model, ref2_model):
        for n, p in model.named_parameters():
            if p.requires_grad:
                p.grad = eng.param_view(n, eng.grads_bf16).clone()
    ours.step(); ref.step(); ref2.step(); ours3.step()
    for i, n in enumerate(ours.names):
        a = eng.param_view(n)
        _same_weights(a, dict(ref_model.named_parameters())[n].detach(), n)
        _same_weights(a, dict(ref2_model.named_parameters())[n].detach(), n)
        _same_weights(a, eng3.param_view(n), n)
        # the fp32 masters (what a resume continues from) agree to fp32 rounding of the clip coefficient
        torch.testing.assert_close(eng.param_view(n, eng.master), ref2.fp32_params[i].detach(), rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(eng.param_view(n, eng.master), eng3.param_view(n, eng3.master), rtol=1e-5, atol=1e-7)
    assert ours.get_learning_rate() == pytest.approx(ref2.get_learning_rate()) == pytest.approx(ours3.get_learning_rate())


def test_optimizer_load_validates_like_the_reference():
    from specforge_b200.optimizer import B200BF16Optimizer
    eng = CpuFlatEngine(SHAPES)
    opt = B200BF16Optimizer(_Holder(eng), lr=1e-3, max_grad_norm=0.5, total_steps=50, warmup_ratio=0.1)
    _random_grads(eng, 0)
    opt.step()
    sd = opt.state_dict()
    other = B200BF16Optimizer(_Holder(CpuFlatEngine(SHAPES)), lr=1e-3, max_grad_norm=1.0, total_steps=50, warmup_ratio=0.1)
    with pytest.raises(ValueError, match="max_grad_norm"):
        other.load_state_dict(sd)
    other = B200BF16Optimizer(_Holder(CpuFlatEngine(SHAPES)), lr=1e-3, max_grad_norm=0.5, total_steps=50, warmup_ratio=0.1,
                              lr_scheduler="constant")
    with pytest.raises(ValueError, match="lr_scheduler"):
        other.load_state_dict(sd)
    bad = copy.deepcopy(sd)
    bad["fp32_params"] = bad["fp32_params"][:-1]
    with pytest.raises(ValueError, match="fp32 master"):
        B200BF16Optimizer(_Holder(CpuFlatEngine(SHAPES)), lr=1e-3, max_grad_norm=0.5, total_steps=50, warmup_ratio=0.1).load_state_dict(bad)
    # save -> load -> step == uninterrupted step, bit for bit
    eng2 = CpuFlatEngine(SHAPES, seed=5)
    eng2.params.copy_(eng.params)
    opt2 = B200BF16Optimizer(_Holder(eng2), lr=1e-3, max_grad_norm=0.5, total_steps=50, warmup_ratio=0.1)
    opt2.load_state_dict(copy.deepcopy(sd))
    _random_grads(eng, 1)
    eng2.grads_bf16.copy_(eng.grads_bf16)
    opt.step(); opt2.step()
    assert torch.equal(eng.params, eng2.params) and torch.equal(eng.master, eng2.master) and torch.equal(eng.exp_avg_sq, eng2.exp_avg_sq)


@needs_ref
def test_registry_and_patch_points():
    from specforge.algorithms.builtin import builtin_algorithm_registry
    from specforge_b200 import train as T
    from specforge_b200.backend import B200TrainingBackend
    from specforge_b200.optimizer import B200BF16Optimizer
    reg = T.b200_registry()
    assert reg.names == builtin_algorithm_registry().names
    ours, stock = reg.resolve("eagle3"), builtin_algorithm_registry().resolve("eagle3")
    assert ours.providers.step.build is T.build_step and ours.providers.model.build_draft is T.build_draft
    assert ours.providers.model.build_training_model is T.build_training_model
    assert ours.spec == stock.spec                                     # contracts / capabilities untouched
    assert ours.providers.offline == stock.providers.offline or [type(o) for o in ours.providers.offline] == [type(o) for o in stock.providers.offline]
    assert reg.resolve("dflash").providers.step.build is builtin_algorithm_registry().resolve("dflash").providers.step.build
    import specforge.optimizer as ref_opt
    import specforge.training.trainer as ref_trainer
    stock_backend, stock_opt = ref_trainer.FSDPTrainingBackend, ref_opt.BF16Optimizer
    with T.b200_patches():
        assert ref_trainer.FSDPTrainingBackend is B200TrainingBackend and ref_opt.BF16Optimizer is B200BF16Optimizer
    assert ref_trainer.FSDPTrainingBackend is stock_backend and ref_opt.BF16Optimizer is stock_opt


@needs_ref
def test_dflash_loss_terms_through_the_reference_trainer_core():
    """The reference TrainerCore (controller.py:328-398) on a strategy that reports loss_terms: it backpropagates the NUMERATOR
    divided by the accumulation steps and, at the boundary, scales the gradients by world * accum / sum(denominators).  With the
    B200 backend + DFlash strategy in normalization="global" the accumulated gradient must come out as sum_i d(num_i) / sum_i den_i."""
    from specforge.training.controller import TrainerCore
    from specforge_b200.backend import B200TrainingBackend
    from specforge_b200.contracts import TrainBatch
    from specforge_b200.dflash import B200DFlashTrainStrategy

    class Eng(CpuFlatEngine):
        """forward: loss_num = c * sum(params), den = d for scripted (c, d); backward adds d(loss_num) = c (numerator mode)."""
        script = [(3.0, 2.0), (5.0, 6.0)]

        def __init__(self):
            super().__init__({"w": (8,)})
            self.i, self.metrics, self.loss, self.grad_of_numerator = 0, torch.zeros(4), torch.zeros(1), 0
            self.names = ["w"]

        def set_frozen(self, **kw):
            pass

        def forward(self, batch, anchors, keep, need_grad=True):
            c, d = self.script[self.i % 2]
            self.cur = c
            num = c * float(self.params.float().sum())
            self.metrics = torch.tensor([num, d, 1.0, 2.0])
            self.loss = torch.tensor([num / d])
            self.i += 1
            return self.loss, self.metrics

        def backward(self, accumulate=False):
            assert self.grad_of_numerator == 1
            g = torch.full((8,), self.cur)
            self.grads_f32 = self.grads_f32 + g if accumulate else g

    class Draft(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.engine = Eng()

    draft = Draft()
    eng = draft.engine
    backend = B200TrainingBackend(lr=1e-2, max_grad_norm=1e9, total_steps=10, warmup_ratio=0.0)
    st = B200DFlashTrainStrategy(draft, target_embed_weight=torch.zeros(4, 4), target_head_weight=torch.zeros(4, 4), num_anchors=3,
                                 generator=torch.Generator().manual_seed(0), normalization="global")
    backend.attach(st)
    backend.prepare_model(st.trainable_module())
    core = TrainerCore(st, backend, accumulation_steps=2)
    batch = TrainBatch(sample_ids=["0"], strategy="dflash", tensors={"input_ids": torch.zeros(1, 12, dtype=torch.long),
                                                                     "hidden_states": torch.zeros(1, 12, 8), "loss_mask": torch.ones(1, 12)}, metadata={})
    r1 = core.train_step(batch)
    assert not r1.optimizer_stepped
    seen = {}
    orig = eng.optimizer_step

    def spy(lr, **kw):
        seen["g"] = eng.grads_bf16.float().clone()
        return orig(lr, **kw)

    eng.optimizer_step = spy
    r2 = core.train_step(batch)
    assert r2.optimizer_stepped
    # (3 + 5) / (2 + 6) = 1.0 per element: sum of numerator gradients over the sum of denominators
    torch.testing.assert_close(seen["g"], torch.full((8,), 1.0), rtol=1e-2, atol=1e-2)
