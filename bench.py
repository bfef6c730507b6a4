#!/usr/bin/env python
"""bench.py — EAGLE3 draft-step samples/sec (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm (CPU port) on the host cores

One "step" = teacher + TTT-unrolled forward + loss + full backward + gradient all-reduce (N>1) + clip/AdamW on one
micro-batch per GPU.  Workload at every N: BASELINE config 2 per GPU — Qwen3-8B EAGLE3 draft (configs/qwen3-8b-eagle3.json
dims), B=8 sequences x S=2048, TTT=7, synthetic hidden states, random-init weights (weak scaling, DP).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

QWEN3_8B = dict(hidden_size=4096, intermediate_size=12288, num_heads=32, num_kv_heads=8, head_dim=128, vocab_size=151936,
                draft_vocab_size=32000, rms_norm_eps=1e-6, rope_theta=1000000.0, max_position_embeddings=40960)
B, S, T = 8, 2048, 7
METRIC = "EAGLE3 draft-step samples/sec (Qwen3-8B, TTT=7, seq 2048)"
WORKLOAD = "BASELINE config 2 per GPU: Qwen3-8B EAGLE3 offline draft step"
# the other EAGLE3 configurations of BASELINE.json (opt-in with --config; parity-test cases, not the headline):
OTHER_CONFIGS = {
    3: (dict(hidden_size=4096, intermediate_size=14336, num_heads=32, num_kv_heads=8, head_dim=128, vocab_size=128256, draft_vocab_size=32000,
             rms_norm_eps=1e-5, rope_theta=10000.0, max_position_embeddings=2048), 8, 2048,
        "EAGLE3 draft-step samples/sec (Llama3-8B, TTT=7, seq 2048)", "BASELINE config 3 per GPU: Llama3-8B EAGLE3 offline draft step"),
    5: (dict(hidden_size=2048, intermediate_size=12288, num_heads=32, num_kv_heads=4, head_dim=128, vocab_size=151936, draft_vocab_size=32000,
             rms_norm_eps=1e-6, rope_theta=1000000.0, max_position_embeddings=8192, fc_norm=True), 4, 4096,
        "EAGLE3.1 draft-step samples/sec (Qwen3-30B-A3B, TTT=7, seq 4096)",
        "BASELINE config 5 per GPU: Qwen3-30B-A3B EAGLE3.1 (fc_norm) offline draft step, max_position_embeddings raised to 8192"),
}
CPU_SAMPLE_TOKENS = 128
CPU_MAX_THREADS = 32


def gemm_traffic():
    """`roofline.traffic`: DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of the CURRENT
    kernel sources (profiles/gemm_traffic.json, written by tools/ncu_traffic.py together with a hash of csrc/sf_gemm*).  A capture
    taken from other sources is reported as stale (traffic: null) instead of being quoted."""
    import hashlib
    out = {"traffic": None, "traffic_algorithmic": None, "traffic_kernel": None}
    try:
        with open(os.path.join(ROOT, "profiles", "gemm_traffic.json")) as f:
            t = json.load(f)
        h = hashlib.sha1()
        for name in ("sf_gemm.cuh", "sf_gemm_wide.cuh", "sf_gemm.cu"):
            with open(os.path.join(ROOT, "specforge_b200", "csrc", name), "rb") as f:
                h.update(f.read())
        if t.get("src_sha1") == h.hexdigest():
            out.update(traffic=t["dram_read_bytes"] + t["dram_write_bytes"], traffic_algorithmic=t["algorithmic_bytes"],
                       traffic_kernel=t["kernel"], traffic_source=t.get("source"))
        else:
            out["traffic_note"] = "profiles/gemm_traffic.json was captured from other kernel sources (stale): not quoted"
    except (OSError, KeyError, ValueError):
        out["traffic_note"] = "no ncu capture committed for this build"
    return out


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["bf16_tflops_sustained"], p["bf16_tflops"], "measured (MEASURED_PEAKS.json, sustained for a long step)"
    except Exception:
        return 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, n in enumerate(names):
                if len(r) > 3 + i and r[3 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(float(self.rows[0][1])) if self.rows else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# --------------------------------------------------------------------------------------------- CPU arm
def _reference_dir():
    d = os.path.join(ROOT, "baseline", "_ref")
    return d if os.path.isdir(os.path.join(d, "specforge")) else None


def reference_cpu_step_rate(steps: int, warmup: int):
    """The UNMODIFIED reference (offline install baseline/_ref) on the host cores, same configuration as the GPU
    arm except the batch: Eagle3TrainStrategy.forward_loss -> TargetHead -> OnlineEagle3Model(sdpa) -> LlamaForCausalLMEagle3 ->
    backward -> BF16Optimizer.step, Qwen3-8B draft dims, B = 1 sequence x S = 2048 tokens, TTT = 7, bf16 modules.  Two shims, the ones
    the reference's own CPU tests use: TORCHDYNAMO_DISABLE=1 and LogSoftmaxLoss -> the reference's torch twin
    _compute_loss (its Triton kernel cannot launch without a GPU).  Returns (samples/s, ms/step, threads)."""
    os.environ["TORCHDYNAMO_DISABLE"] = "1"
    ref = _reference_dir()
    sys.path.insert(0, ref)
    import tempfile
    import torch
    cores = min(os.cpu_count() or 1, CPU_MAX_THREADS)
    torch.set_num_threads(cores)
    import specforge.algorithms.eagle3.model as ref_model
    from specforge.core.loss import _compute_loss

    class _TorchLogSoftmaxLoss:
        @staticmethod
        def apply(logits, target_p, position_mask):
            return _compute_loss(logits, target_p, position_mask)

    ref_model.LogSoftmaxLoss = _TorchLogSoftmaxLoss
    from transformers import LlamaConfig
    from specforge.algorithms.eagle3.model import OnlineEagle3Model
    from specforge.modeling.draft.llama3_eagle import LlamaForCausalLMEagle3
    from specforge.modeling.target.target_head import TargetHead
    from specforge.optimizer import BF16Optimizer
    from specforge.runtime.contracts import TrainBatch
    from specforge.training.strategies.base import Eagle3TrainStrategy
    c = QWEN3_8B
    hf = LlamaConfig(hidden_size=c["hidden_size"], intermediate_size=c["intermediate_size"], num_attention_heads=c["num_heads"],
                     num_key_value_heads=c["num_kv_heads"], num_hidden_layers=1, vocab_size=c["vocab_size"], rms_norm_eps=c["rms_norm_eps"],
                     max_position_embeddings=c["max_position_embeddings"], hidden_act="silu", tie_word_embeddings=False, pad_token_id=0,
                     rope_theta=c["rope_theta"])
    hf.head_dim, hf.draft_vocab_size, hf.rope_theta = c["head_dim"], c["draft_vocab_size"], c["rope_theta"]
    torch.manual_seed(0)
    draft = LlamaForCausalLMEagle3(hf, attention_backend="sdpa")
    g = torch.Generator().manual_seed(0)
    V, H, DV = c["vocab_size"], c["hidden_size"], c["draft_vocab_size"]
    ids = torch.randperm(V, generator=g)[:DV].sort().values
    draft.t2d.zero_()
    draft.t2d[ids] = True
    draft.d2t.copy_(ids - torch.arange(DV))
    draft = draft.to(torch.bfloat16)
    draft.freeze_embedding()
    model = OnlineEagle3Model(draft_model=draft, length=T, attention_backend="sdpa")
    with tempfile.TemporaryDirectory() as wd:
        with open(os.path.join(wd, "config.json"), "w") as f:
            json.dump({"architectures": ["LlamaForCausalLM"], "model_type": "llama", "hidden_size": H, "vocab_size": V,
                       "num_hidden_layers": 1, "num_attention_heads": 4, "intermediate_size": 128}, f)
        head = TargetHead(wd)
    with torch.no_grad():
        head.fc.weight.copy_(torch.randn(V, H, generator=g))
    head.freeze_weights()
    head = head.eval().to(torch.bfloat16)
    strategy = Eagle3TrainStrategy(model, target_head=head, ploss_decay=0.8)
    opt = BF16Optimizer(draft, lr=1e-4, max_grad_norm=0.5, total_steps=100000, warmup_ratio=0.015)
    times = []
    for it in range(warmup + steps):
        gi = torch.Generator().manual_seed(100 + it)
        t = {"input_ids": torch.randint(0, V, (1, S), generator=gi), "attention_mask": torch.ones(1, S, dtype=torch.long),
             "loss_mask": torch.ones(1, S, dtype=torch.long), "hidden_state": torch.randn(1, S, 3 * H, generator=gi).bfloat16(),
             "target": torch.randn(1, S, H, generator=gi).bfloat16()}
        t["loss_mask"][:, -1] = 0
        tb = TrainBatch(sample_ids=["0"], strategy="eagle3", tensors=t, metadata={"target_repr": "hidden_state"})
        t0 = time.perf_counter()
        out = strategy.forward_loss(tb)
        out.loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    return 1.0 / (ms / 1e3), ms, cores


def cpu_reference_step_rate(steps: int, warmup: int, tokens: int = CPU_SAMPLE_TOKENS):
    """Fallback when the reference install is absent: the oracle port of the PyTorch path (bf16 modules, torch autograd + AdamW on
    fp32 masters) on the host cores.  Bounded sample: 1 sequence x `tokens` tokens of the config-2 dims, TTT=7; samples/s is scaled by
    tokens (tokens/2048 of a sample per step)."""
    import torch
    from oracle import eagle3_oracle as O
    cores = min(os.cpu_count() or 1, CPU_MAX_THREADS)   # more threads only add oneDNN/OpenMP contention on this path
    torch.set_num_threads(cores)
    cfg = O.Eagle3Config(ttt_length=T, **{k: v for k, v in QWEN3_8B.items()})
    g = torch.Generator().manual_seed(0)
    P = {}
    for name, shape in O.param_shapes(cfg).items():
        P[name] = (torch.ones(shape) if len(shape) == 1 else torch.empty(shape).uniform_(-0.03, 0.03, generator=g)).to(torch.bfloat16)
    P["embed_tokens.weight"] = torch.empty(cfg.vocab_size, cfg.hidden_size).uniform_(-0.03, 0.03, generator=g).to(torch.bfloat16)
    head_w = torch.empty(cfg.vocab_size, cfg.target_hidden_size).uniform_(-1, 1, generator=g).to(torch.bfloat16)
    t2d, d2t = O.make_vocab_map(cfg.vocab_size, cfg.draft_vocab_size, seed=0)
    names = [n for n in O.PARAM_NAMES]
    params = [P[n] for n in names]
    masters = [p.float() for p in params]
    ea = [torch.zeros_like(m) for m in masters]
    es = [torch.zeros_like(m) for m in masters]
    times = []
    for it in range(warmup + steps):
        batch = O.make_batch(cfg, 1, tokens, seed=it)
        t0 = time.perf_counter()
        res, grads = O.train_step(P, cfg, batch, head_w, t2d, d2t)
        O.adamw_clip_step(params, masters, ea, es, [grads[n] for n in names], step=it + 1, lr=1e-4)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    return (tokens / S) / (ms / 1e3), ms, cores


def cpu_arm(steps: int, warmup: int):
    """(value, ms, cores, kind, sample): the unmodified reference when its install travelled with the repo, else the port."""
    if _reference_dir() is not None:
        val, ms, cores = reference_cpu_step_rate(steps, warmup)
        return val, ms, cores, "reference", (f"UNMODIFIED reference (baseline/_ref: Eagle3TrainStrategy + OnlineEagle3Model(sdpa) + BF16Optimizer), "
                                             f"Qwen3-8B draft dims, batch 1 x {S} tokens, TTT={T}, fwd+bwd+optimizer, bf16 modules; "
                                             f"{steps} timed step(s) after {warmup} warm-up; one step = one sample")
    val, ms, cores = cpu_reference_step_rate(steps, warmup)
    return val, ms, cores, "port", (f"oracle port (reference install absent): 1 sequence x {CPU_SAMPLE_TOKENS} tokens of the Qwen3-8B draft dims, "
                                    f"TTT={T}, fwd+bwd+AdamW, scaled by tokens; {steps} timed step(s) after {warmup} warm-up")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 2))          # one reference step at these dims is tens of seconds on the host cores
    warm = max(0, min(args.warmup, 1))
    val, ms, cores, kind, sample = cpu_arm(steps, warm)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "BASELINE config 2: Qwen3-8B EAGLE3 offline draft step, TTT=7, seq 2048 (CPU arm: batch 1 per step)",
                       "ttt_length": T, "seq_len": S, "batch_per_step": 1, "same_config": kind == "reference"},
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- GPU arm
def kernel_timeline(step_fn, out_path, steps=2):
    """CUPTI kernel timeline of `steps` steps (torch.profiler, CUDA activities only): per-stream busy time, idle gaps on
    the busiest stream, time per kernel name.  Diagnostic only — nothing measured under the profiler is a bench value."""
    import gzip, tempfile, collections
    import torch
    from torch.profiler import profile, ProfilerActivity
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            step_fn()
        torch.cuda.synchronize()
    tmp = tempfile.mktemp(suffix=".json")
    prof.export_chrome_trace(tmp)
    ev = [e for e in json.load(open(tmp))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    os.remove(tmp)
    ev.sort(key=lambda e: e["ts"])
    t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
    streams = collections.defaultdict(list)
    for e in ev:
        streams[e["args"].get("stream")].append(e)
    main = max(streams, key=lambda k: sum(e["dur"] for e in streams[k]))
    by_name = collections.defaultdict(lambda: [0, 0.0])
    for e in ev:
        k = (e["name"][:90], "main" if e["args"].get("stream") == main else "side")
        by_name[k][0] += 1
        by_name[k][1] += e["dur"]
    gaps = []
    m = streams[main]
    for a, b in zip(m, m[1:]):
        g = b["ts"] - (a["ts"] + a["dur"])
        if g > 3.0:
            gaps.append({"gap_us": round(g, 1), "after": a["name"][:60], "before": b["name"][:60]})
    # union busy time over all streams
    busy, cur_s, cur_e = 0.0, None, None
    for e in ev:
        s, en = e["ts"], e["ts"] + e["dur"]
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                busy += cur_e - cur_s
            cur_s, cur_e = s, en
        else:
            cur_e = max(cur_e, en)
    busy += cur_e - cur_s
    out = {
        "steps": steps, "span_ms": (t1 - t0) / 1e3, "any_stream_busy_ms": busy / 1e3, "idle_ms": (t1 - t0 - busy) / 1e3,
        "streams": {str(k): {"kernels": len(v), "busy_ms": sum(e["dur"] for e in v) / 1e3} for k, v in streams.items()},
        "main_stream": str(main), "main_gaps_over_3us": len(gaps), "main_gap_total_ms": sum(g["gap_us"] for g in gaps) / 1e3,
        "largest_gaps": sorted(gaps, key=lambda g: -g["gap_us"])[:40],
        "by_kernel_ms": [{"name": k[0], "stream": k[1], "n": v[0], "ms": round(v[1] / 1e3, 3)}
                         for k, v in sorted(by_name.items(), key=lambda kv: -kv[1][1])],
    }
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    with open(out_path, "w") as f:
        json.dump(out, f, indent=1)
    with gzip.open(out_path + ".events.gz", "wt") as f:
        json.dump([[e["name"][:60], e["args"].get("stream"), e["ts"] - t0, e["dur"]] for e in ev], f)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from specforge_b200._lib import lib
    from specforge_b200.backend import B200TrainingBackend
    from specforge_b200.contracts import TrainBatch
    from specforge_b200.draft import B200Eagle3DraftModel
    from specforge_b200.strategy import B200Eagle3TrainStrategy

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the EAGLE3 hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = lib()
    import ctypes
    L.sf_profile_gemm.restype = None
    L.sf_profile_gemm_collect.restype = ctypes.c_longlong
    L.sf_profile_gemm_collect.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]

    cfg = dict(QWEN3_8B)
    cfg.update(num_attention_heads=cfg.pop("num_heads"), num_key_value_heads=cfg.pop("num_kv_heads"))
    draft = B200Eagle3DraftModel(cfg)
    eng = draft.bind_engine(batch=B, seq_len=S, ttt_length=T, device=dev, seed=0)
    # frozen tables: random (no checkpoints offline), identical on every rank
    gen = torch.Generator(device=dev).manual_seed(0)
    V, H, DV = cfg["vocab_size"], cfg["hidden_size"], cfg["draft_vocab_size"]
    draft.embed_tokens_weight.data = (torch.randn(V, H, device=dev, generator=gen) * 0.02).bfloat16()
    head_w = torch.randn(V, H, device=dev, generator=gen)
    perm = torch.randperm(V, device=dev, generator=gen)[:DV].sort().values
    # a real draft vocabulary holds the frequent tokens, so the teacher's argmax lands inside it for ~95 % of positions (the
    # rows whose loss is live); with i.i.d. rows it would be DV/V = 21 %.  Doubling the draft-vocab rows of the random head
    # reproduces that coverage, so the loss kernel's second pass runs on a realistic share of the rows.
    head_w[perm] *= 2.0
    head_w = head_w.bfloat16()
    draft.t2d.zero_()
    draft.t2d[perm] = True
    draft.d2t.copy_(perm - torch.arange(DV, device=dev))
    strategy = B200Eagle3TrainStrategy(draft, target_head_weight=head_w)
    backend = B200TrainingBackend(lr=1e-4, max_grad_norm=0.5, total_steps=100000, warmup_ratio=0.015)
    backend.attach(strategy)
    backend.prepare_model(strategy.trainable_module())

    # synthetic micro-batch, per-rank data seed
    dgen = torch.Generator(device=dev).manual_seed(1000 + rank)
    dev_t = {
        "input_ids": torch.randint(0, V, (B, S), device=dev, generator=dgen),
        "attention_mask": torch.ones(B, S, dtype=torch.long, device=dev),
        "loss_mask": torch.ones(B, S, dtype=torch.long, device=dev),
        "hidden_state": torch.randn(B, S, 3 * H, device=dev, generator=dgen).bfloat16(),
        "target": torch.randn(B, S, H, device=dev, generator=dgen).bfloat16(),
    }
    dev_t["loss_mask"][:, -1] = 0
    host_t = {k: v.cpu().pin_memory() for k, v in dev_t.items()}
    h2d = sum(v.numel() * v.element_size() for v in host_t.values())
    ids = [str(i) for i in range(B)]
    dev_batch = TrainBatch(sample_ids=ids, strategy="eagle3", tensors=dev_t, metadata={"target_repr": "hidden_state"})
    host_batch = TrainBatch(sample_ids=ids, strategy="eagle3", tensors=host_t, metadata={"target_repr": "hidden_state"})

    def step(batch, read_loss):
        out = strategy.forward_loss(batch)
        backend.backward(out.loss, is_boundary=True)
        backend.step()
        return float(out.loss.item()) if read_loss else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batch, read_loss, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            last = step(batch, read_loss)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, last

    for _ in range(max(3, args.warmup)):
        step(dev_batch, False)
    torch.cuda.synchronize()
    # ---- kernel-resident measurement ("value"): inputs already in HBM
    L.sf_launch_count_reset()
    sampler = ClockSampler(local)
    sampler.start()
    ms_total, _ = timed(dev_batch, False, args.steps)
    clocks = sampler.stop()
    launches = int(L.sf_launch_count())
    # ---- roofline pass: the same K steps again with every GEMM launch bracketed by CUDA events on its stream (kept out of
    # the `value` region: the events sit between launches and switch off the tail overlap of the weight-gradient GEMMs)
    L.sf_profile_gemm(1)
    ms_prof, _ = timed(dev_batch, False, args.steps)
    gms, gfl = ctypes.c_double(), ctypes.c_double()
    n_gemm = int(L.sf_profile_gemm_collect(ctypes.byref(gms), ctypes.byref(gfl)))
    # the same launches by shape: which GEMMs run at what rate inside the step (live, power-capped)
    L.sf_profile_gemm_detail.restype = ctypes.c_longlong
    cap = max(1, n_gemm)
    mnkt, tms = (ctypes.c_longlong * (4 * cap))(), (ctypes.c_double * cap)()
    nd = int(L.sf_profile_gemm_detail(mnkt, tms, cap))
    by_shape = {}
    for i in range(nd):
        key = tuple(mnkt[4 * i + k] for k in range(4))
        c = by_shape.setdefault(key, [0, 0.0])
        c[0] += 1
        c[1] += tms[i]
    gemm_by_shape = [{"M": k[0], "N": k[1], "K": k[2], "tile_rows": k[3], "launches_per_step": v[0] / args.steps,
                      "ms_per_step": round(v[1] / args.steps, 3), "tflops": round(2.0 * k[0] * k[1] * k[2] * v[0] / v[1] / 1e9, 1)}
                     for k, v in sorted(by_shape.items(), key=lambda kv: -kv[1][1])]
    L.sf_profile_gemm(0)
    ms_step = ms_total / args.steps
    value = world * B / (ms_step / 1e3)
    # ---- end-to-end through the public API with HOST (pinned) inputs: every step's batch crosses PCIe inside the timed
    # region (double-buffered on a side stream by DevicePrefetcher) and the loss is read back to the host every step
    from specforge_b200.feed import DevicePrefetcher
    for batch in DevicePrefetcher((host_batch for _ in range(3)), device=dev):   # warm the feed path as well
        step(batch, True)

    def timed_e2e(n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for batch in DevicePrefetcher((host_batch for _ in range(n)), device=dev):
            last = step(batch, True)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, last

    ms_e2e, last_loss = timed_e2e(args.steps)
    e2e = world * B / (ms_e2e / args.steps / 1e3)
    if args.e2e_variants and rank == 0:
        # diagnostic: which part of the end-to-end loop costs what (device batch + loss read; host feed without the read)
        ms_b, _ = timed(dev_batch, True, args.steps)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for batch in DevicePrefetcher((host_batch for _ in range(args.steps)), device=dev):
            step(batch, False)
        e1.record()
        barrier()
        print(json.dumps({"e2e_variants_ms_per_step": {"resident_no_read": ms_total / args.steps, "resident_loss_read": ms_b / args.steps,
                                                       "host_feed_no_read": e0.elapsed_time(e1) / args.steps, "host_feed_loss_read": ms_e2e / args.steps}}),
              file=sys.stderr, flush=True)

    e2e_shards = None
    if args.feed_shards:
        # optional: the same end-to-end loop fed from an SFPK shard on local disk (specforge_b200/shards.py): every step's
        # records are gathered from the file into pinned batch tensors by the C++ reader, then H2D, then the step
        import tempfile
        from specforge_b200.shards import Eagle3ShardLoader, ShardWriter
        tmpd = tempfile.mkdtemp(prefix="sfpk_bench_")
        shard = os.path.join(tmpd, f"rank{rank}.sfpk")
        feats = [("input_ids", torch.int64, 1), ("loss_mask", torch.int64, 1), ("hidden_state", torch.bfloat16, H),
                 ("aux_hidden_state", torch.bfloat16, 3 * H)]
        with ShardWriter(shard, feats) as w:
            for rep in range(2):
                for i in range(B):
                    w.add({"input_ids": host_t["input_ids"][i], "loss_mask": host_t["loss_mask"][i],
                           "hidden_state": host_t["target"][i], "aux_hidden_state": host_t["hidden_state"][i]})
        loader = Eagle3ShardLoader([shard], batch_size=B, max_len=S, pad_to=S, threads=8, buffers=4)

        def shard_batches(n):
            done = 0
            while done < n:
                for b in loader:
                    yield b
                    done += 1
                    if done == n:
                        return

        for _ in shard_batches(8):          # allocate (pin) every buffer set of the loader's ring before timing
            pass
        for batch in DevicePrefetcher(shard_batches(3), device=dev):
            step(batch, True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for batch in DevicePrefetcher(shard_batches(args.steps), device=dev):
            last_s = step(batch, True)
        e1.record()
        barrier()
        ms_sh = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms_sh], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_sh = float(t.item())
        e2e_shards = {"value": world * B / (ms_sh / args.steps / 1e3), "unit": "samples/s", "ms_per_step": ms_sh / args.steps,
                      "disk_read_bytes_per_step": h2d, "last_loss": last_s, "source": "SFPK shard on local disk, page cache warm"}
        import shutil
        shutil.rmtree(tmpd, ignore_errors=True)

    for ab in (args.ab or []) if rank == 0 else []:
        # in-process A/B of a library switch: the settings alternate step by step, so box-to-box clock differences cancel
        import statistics
        name, vals = ab.split("=")
        vals = [int(v) for v in vals.split(",")]
        acc = {v: [] for v in vals}
        for rnd in range(args.ab_rounds + 1):
            for v in vals:
                if L.sf_debug_option(name.encode(), v) != 0:
                    raise SystemExit("unknown option " + name)
                ms, _ = timed(dev_batch, False, 1)
                if rnd:                      # round 0 warms each setting
                    acc[v].append(ms)
        L.sf_debug_option(name.encode(), vals[0])
        print(json.dumps({"ab": name, "rounds": args.ab_rounds,
                          "ms_per_step": {str(v): {"mean": round(statistics.mean(t), 3), "median": round(statistics.median(t), 3),
                                                   "min": round(min(t), 3)} for v, t in acc.items()}}), file=sys.stderr, flush=True)
    if args.timeline and rank == 0:
        kernel_timeline(lambda: step(dev_batch, False), args.timeline)

    sustained, burst, peak_src = peaks()
    flops_step = eng.flops_per_step()
    achieved = (gfl.value / 1e12) / (gms.value / 1e3) if gms.value > 0 else 0.0
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD + " (teacher + TTT fwd + loss + bwd + grad all-reduce + clip/AdamW)", "batch_per_gpu": B, "global_batch": world * B, "seq_len": S,
                   "ttt_length": T, "parallelism": f"dp{world}", "l2": "inputs_exceed_l2 (per-step working set ~45 GB)",
                   "weights": "random-init, reference shapes; draft-vocab rows of the frozen head x2 so ~95 % of positions are live"},
        "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "last_loss": last_loss},
        "e2e_shards": e2e_shards,
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "sf::gemm_kernel (tcgen05, all %d launches/step)" % (n_gemm // max(1, args.steps)),
                     "achieved": achieved, "peak": sustained, "unit": "TFLOP/s", "frac": achieved / sustained if sustained else None,
                     "peak_burst": burst, "frac_of_burst": achieved / burst if burst else None, "peak_source": peak_src,
                     "gemm_ms_per_step": gms.value / args.steps, "gemm_share_of_step": (gms.value / args.steps) / (ms_prof / args.steps),
                     "ms_per_step_with_events": ms_prof / args.steps, "gemm_by_shape": gemm_by_shape,
                     "step_tflops_algorithmic": flops_step / 1e12 / (ms_step / 1e3),
                     "step_frac_of_burst": flops_step / 1e12 / (ms_step / 1e3) / burst if burst else None,
                     **gemm_traffic()},
    }
    if world > 1:
        # driver-visible DP correctness: after K optimizer steps on different data every replica must hold the same weights
        # (one all-reduced gradient, identical fused updates) and have seen the same global gradient norm
        chk = torch.stack([eng.params.float().sum(), eng.params.float().square().sum(), eng.grad_norm.reshape(()).float()]).double()
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        line["replicas_identical"] = bool(all(torch.equal(allc[0], c) for c in allc))
        line["replica_checksums"] = [[float(v) for v in c.tolist()] for c in allc] if not line["replicas_identical"] else [float(v) for v in allc[0].tolist()]
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            val, ms, cores, kind, sample = cpu_arm(1, 0)
            line["cpu_baseline"] = {"value": val, "unit": "samples/s", "cores": cores, "kind": kind, "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_dflash(args):
    """BASELINE config 4 (next row, NOT the headline metric): DFlash draft step, Qwen3-8B dims, 4 sequences x 2048 tokens per
    GPU, 512 anchors x block 16, full-vocabulary frozen head; fwd + loss + bwd + all-reduce + clip/AdamW through
    B200DFlashTrainStrategy / B200TrainingBackend.  Same JSON line layout as the EAGLE3 workload."""
    import ctypes
    import torch
    import torch.distributed as dist
    from specforge_b200._lib import lib
    from specforge_b200.backend import B200TrainingBackend
    from specforge_b200.contracts import TrainBatch
    from specforge_b200.dflash import B200DFlashDraftModel, B200DFlashTrainStrategy
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = lib()
    L.sf_profile_gemm.restype = None
    L.sf_profile_gemm_collect.restype = ctypes.c_longlong
    L.sf_profile_gemm_collect.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    Bd, Sd, Nd, bs, H, V, F = 4, 2048, 512, 16, 4096, 151936, 5
    cfg = dict(hidden_size=H, intermediate_size=12288, num_attention_heads=32, num_key_value_heads=8, head_dim=128, num_hidden_layers=5,
               vocab_size=V, block_size=bs, rms_norm_eps=1e-6, rope_theta=1000000.0, max_position_embeddings=40960,
               dflash_config={"mask_token_id": 151669, "target_layer_ids": [1, 9, 17, 25, 33]}, layer_types=["full_attention"] * 5)
    draft = B200DFlashDraftModel(cfg)
    eng = draft.bind_engine(batch=Bd, seq_len=Sd, num_anchors=Nd, device=dev, seed=0)
    gen = torch.Generator(device=dev).manual_seed(0)
    embed = (torch.randn(V, H, device=dev, generator=gen) * 0.02).bfloat16()
    head = (torch.randn(V, H, device=dev, generator=gen) * 0.02).bfloat16()
    strategy = B200DFlashTrainStrategy(draft, target_embed_weight=embed, target_head_weight=head, num_anchors=Nd,
                                       generator=torch.Generator(device=dev).manual_seed(100 + rank))
    backend = B200TrainingBackend(lr=1e-4, max_grad_norm=1.0, total_steps=100000, warmup_ratio=0.015)
    backend.attach(strategy)
    backend.prepare_model(strategy.trainable_module())
    dgen = torch.Generator(device=dev).manual_seed(1000 + rank)
    dev_t = {"input_ids": torch.randint(0, V - 1000, (Bd, Sd), device=dev, generator=dgen),
             "hidden_states": torch.randn(Bd, Sd, F * H, device=dev, generator=dgen).bfloat16(),
             "loss_mask": torch.ones(Bd, Sd, device=dev)}
    host_t = {k: v.cpu().pin_memory() for k, v in dev_t.items()}
    h2d = sum(v.numel() * v.element_size() for v in host_t.values())
    ids = [str(i) for i in range(Bd)]
    mk = lambda t: TrainBatch(sample_ids=ids, strategy="dflash", tensors=t, metadata={})

    def step(batch, read_loss):
        out = strategy.forward_loss(batch)
        backend.backward(out.loss, is_boundary=True)
        backend.step()
        return float(out.loss.item()) if read_loss else None

    def timed(batches, read_loss):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last, n = None, 0
        for b in batches:
            last = step(b, read_loss)
            n += 1
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / max(1, n), last

    from specforge_b200.feed import DevicePrefetcher
    for _ in range(max(3, args.warmup)):
        step(mk(dev_t), False)
    L.sf_launch_count_reset()
    sampler = ClockSampler(local)
    sampler.start()
    ms_step, _ = timed((mk(dev_t) for _ in range(args.steps)), False)
    clocks = sampler.stop()
    launches = int(L.sf_launch_count())
    L.sf_profile_gemm(1)
    ms_prof, _ = timed((mk(dev_t) for _ in range(args.steps)), False)
    gms, gfl = ctypes.c_double(), ctypes.c_double()
    L.sf_profile_gemm_collect(ctypes.byref(gms), ctypes.byref(gfl))
    L.sf_profile_gemm(0)
    for b in DevicePrefetcher((mk(host_t) for _ in range(3)), device=dev):
        step(b, True)
    ms_e2e, last_loss = timed(DevicePrefetcher((mk(host_t) for _ in range(args.steps)), device=dev), True)
    if args.timeline and rank == 0:
        kernel_timeline(lambda: step(mk(dev_t), False), args.timeline)
    sustained, burst, peak_src = peaks()
    A, KV, I, Lr, Mq, Mc = 4096, 1024, 12288, 5, Bd * Nd * bs, Bd * Sd
    fwd = 2 * Mc * F * H * H + Lr * (2 * Mq * H * (A + 2 * KV) + 2 * Mc * H * 2 * KV + 2 * Mq * A * H + 6 * Mq * H * I) + 2 * Mq * H * V
    flops = 3 * fwd - 2 * Mq * H * V + 3.5 * Lr * 4 * Bd * 32 * 128 * Nd * bs * (Sd / 2 + bs)     # no wgrad for the frozen head
    achieved = (gfl.value / 1e12) / (gms.value / 1e3) if gms.value > 0 else 0.0
    line = {"metric": "DFlash draft-step samples/sec (Qwen3-8B, 5 layers, block 16, 512 anchors, seq 2048)", "value": world * Bd / (ms_step / 1e3),
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "BASELINE config 4 per GPU: Qwen3-8B DFlash draft step (fwd + loss + bwd + grad all-reduce + clip/AdamW)",
                       "batch_per_gpu": Bd, "global_batch": world * Bd, "seq_len": Sd, "block_size": bs, "num_anchors": Nd,
                       "parallelism": f"dp{world}", "l2": "inputs_exceed_l2", "attention": "cuda-core (A/B baseline)" if os.environ.get("SF_DFLASH_ATTN_TC") == "-1" else "tcgen05"},
            "e2e": {"value": world * Bd / (ms_e2e / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e, "last_loss": last_loss},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "sf::gemm_kernel (tcgen05)", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                         "frac": achieved / sustained if sustained else None, "peak_burst": burst, "peak_source": peak_src,
                         "gemm_ms_per_step": gms.value / args.steps, "ms_per_step_with_events": ms_prof,
                         "step_tflops_algorithmic": flops / 1e12 / (ms_step / 1e3),
                         "step_frac_of_burst": flops / 1e12 / (ms_step / 1e3) / burst if burst else None, "traffic": None}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    wd = int(os.environ.get("SF_BENCH_WATCHDOG", "0"))
    if wd > 0:   # diagnostic: dump every thread's Python stack and exit if the run takes longer than `wd` seconds
        import faulthandler
        faulthandler.dump_traceback_later(wd, exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--ab", action="append", default=None, help="diagnostic: NAME=V0,V1[,..] alternates sf_debug_option(NAME) step by step, prints ms/step per value to stderr")
    ap.add_argument("--ab-rounds", type=int, default=10)
    ap.add_argument("--e2e-variants", action="store_true", help="diagnostic: time the end-to-end loop with / without the host feed and the loss read")
    ap.add_argument("--feed-shards", action="store_true", help="also measure e2e fed from an SFPK shard on local disk")
    ap.add_argument("--timeline", default=None, help="also write a CUPTI kernel-timeline summary of 2 steps to this path")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="eagle3", choices=["eagle3", "dflash"],
                    help="eagle3 = the headline metric (BASELINE config 2); dflash = BASELINE config 4 (next row, opt-in)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 5], help="EAGLE3 BASELINE configuration (2 = the headline metric)")
    args = ap.parse_args()
    if args.config != 2:
        global QWEN3_8B, B, S, METRIC, WORKLOAD
        QWEN3_8B, B, S, METRIC, WORKLOAD = OTHER_CONFIGS[args.config]
        args.no_cpu_baseline = True
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "dflash":
        run_dflash(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
