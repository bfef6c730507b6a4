"""Host-side DP logic of B200TrainingBackend on CPU with gloo, world_size 2 (no GPU, no CUDA library calls):
replica broadcast at prepare_model, ONE all-reduce(sum) of the flat bf16 gradient buffer per optimizer step with the
1/world averaging folded into the optimizer's grad_scale, identical updates on every rank, LR schedule parity."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _FakeEngine:
    """Stands in for Eagle3Engine: same attributes/methods the backend touches, plain SGD arithmetic on CPU."""

    def __init__(self, n, rank):
        self.device = torch.device("cpu")
        self.n_params = n
        self.params = torch.full((n,), float(rank + 1), dtype=torch.bfloat16)   # ranks start DIFFERENT on purpose
        self.grads_f32 = torch.zeros(n)
        self.grads_bf16 = torch.zeros(n, dtype=torch.bfloat16)
        self.master = self.exp_avg = self.exp_avg_sq = None
        self.opt_step = 0
        self.calls = []
        self.offsets, self.names = {}, []          # no named parameters: the backend only touches the flat buffers here

    def grads_to_bf16(self, scale=None):
        s = 1.0 if scale is None else float(scale)
        self.grads_bf16.copy_((self.grads_f32 * s).to(torch.bfloat16))
        return self.grads_bf16

    def optimizer_step(self, lr, *, grad_scale=1.0, max_grad_norm=0.5, weight_decay=0.0, **_):
        self.opt_step += 1
        g = self.grads_bf16.float() * grad_scale
        self.calls.append((lr, grad_scale))
        self.params.copy_((self.params.float() - lr * g).to(torch.bfloat16))
        return g.norm().reshape(1)


class _FakeStrategy:
    def __init__(self, engine):
        self.engine = engine
        self.draft_model = torch.nn.Module()
        self.draft_model.engine = engine
        self._micro_in_window = 0
        self._last_grad_out = torch.tensor(0.5)     # 1/accumulation_steps as autograd would deliver it


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from specforge_b200.backend import B200TrainingBackend
    eng = _FakeEngine(16, rank)
    st = _FakeStrategy(eng)
    be = B200TrainingBackend(lr=0.1, total_steps=100, warmup_ratio=0.1, lr_scheduler="cosine")
    be.attach(st)
    be.prepare_model(torch.nn.Linear(1, 1))
    after_bcast = eng.params.clone()
    # two accumulated micro-batches worth of gradient, different per rank
    eng.grads_f32 += float(rank + 1)
    eng.grads_f32 += float(rank + 1)
    gn = be.step()
    out.put((rank, after_bcast.float().tolist(), eng.params.float().tolist(), eng.calls, float(gn)))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_dp_allreduce_semantics_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (r0, b0, p0, c0, g0), (r1, b1, p1, c1, g1) = res
    assert b0 == b1 == [1.0] * 16                       # rank 0's weights broadcast to every replica
    # accumulated grads: rank0 2.0, rank1 4.0; x0.5 (1/accum) -> 1.0, 2.0; all-reduce sum 3.0; x(1/world) -> 1.5
    lr0 = 0.1 * 1 / 10                                  # warm-up: base_lr * (step+1)/warmup_steps
    assert c0 == c1 == [(pytest.approx(lr0), 0.5)]
    expect = 1.0 - lr0 * 1.5
    assert p0 == p1 and abs(p0[0] - expect) < 4e-3      # identical replicas after the step (bf16 rounding)
    assert g0 == pytest.approx(g1) and g0 == pytest.approx(1.5 * 4.0, rel=1e-2)


def test_lr_schedule_matches_reference_formula():
    from specforge_b200.backend import LRSchedule
    s = LRSchedule(base_lr=1e-4, total_steps=1000, warmup_ratio=0.015, kind="cosine")
    warm = int(0.015 * 1000)
    assert s.lr_at(0) == pytest.approx(1e-4 / warm)
    assert s.lr_at(warm - 1) == pytest.approx(1e-4)
    mid = warm + (1000 - warm) // 2
    # the reference's nested torch CosineAnnealingLR is entered through its chainable form, which scales the cosine phase by
    # 2 / (1 + cos(pi / T_max)) (optimizer.WarmupSchedule; pinned against the reference scheduler in tests/test_train_entry.py)
    tmax = 1000 - warm
    assert s.lr_at(mid) == pytest.approx(1e-4 * (1 + math.cos(math.pi * (mid - warm) / tmax)) / (1 + math.cos(math.pi / tmax)))
    assert LRSchedule(1e-4, 1000, 0.015, "constant").lr_at(500) == 1e-4


def test_backend_exposes_what_the_reference_controller_reads():
    """training/controller.py reads backend.parallel_config.fsdp_process_group, backend.optimizer.get_learning_rate() and
    backend.optimizer_state_is_replicated."""
    from specforge_b200.backend import B200TrainingBackend
    b = B200TrainingBackend(lr=1e-3, total_steps=1000, warmup_ratio=0.1)
    pc = b.parallel_config
    assert pc.world_size == 1 and pc.tp_size == 1 and pc.sharding_strategy == "NO_SHARD" and pc.fsdp_process_group is None
    assert b.optimizer_state_is_replicated is True
    assert b.get_learning_rate() == pytest.approx(1e-3 / 100)      # before attach(); afterwards b.optimizer.get_learning_rate()
