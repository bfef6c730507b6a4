"""The UNMODIFIED reference (baseline/_ref) on the same B200: its own EAGLE3 training step — Eagle3TrainStrategy.forward_loss ->
TargetHead -> OnlineEagle3Model(attention_backend=...) -> LlamaForCausalLMEagle3 (real Triton LogSoftmaxLoss) -> backward ->
BF16Optimizer.step — at the BASELINE config-2 dims (Qwen3-8B draft, S = 2048, TTT = 7), CUDA-event timed.  The "honest competitor"
of BASELINE.md section 3 / SURVEY section 8d: not the optimisation target, a same-box reference point for `bench.py`'s number.

    python tools/reference_gpu_bench.py [--backend flex_attention|sdpa|fa] [--batch 8] [--steps 3] [--warmup 2]

Each backend runs in this process; a backend that cannot run here (Inductor / Triton / flash-attn availability) reports its error.
One JSON line per backend."""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
sys.path.insert(0, REF)


def run(backend, B, steps, warmup):
    import torch
    from transformers import LlamaConfig
    from specforge.algorithms.eagle3.model import OnlineEagle3Model
    from specforge.modeling.draft.llama3_eagle import LlamaForCausalLMEagle3
    from specforge.modeling.target.target_head import TargetHead
    from specforge.optimizer import BF16Optimizer
    from specforge.runtime.contracts import TrainBatch
    from specforge.training.strategies.base import Eagle3TrainStrategy
    dev = torch.device("cuda", 0)
    S, T, H, I, V, DV = 2048, 7, 4096, 12288, 151936, 32000
    hf = LlamaConfig(hidden_size=H, intermediate_size=I, num_attention_heads=32, num_key_value_heads=8, num_hidden_layers=1, vocab_size=V,
                     rms_norm_eps=1e-6, max_position_embeddings=40960, hidden_act="silu", tie_word_embeddings=False, pad_token_id=0,
                     rope_theta=1000000.0)
    hf.head_dim, hf.draft_vocab_size, hf.rope_theta = 128, DV, 1000000.0
    torch.manual_seed(0)
    draft = LlamaForCausalLMEagle3(hf, attention_backend=backend)
    g = torch.Generator().manual_seed(0)
    ids = torch.randperm(V, generator=g)[:DV].sort().values
    draft.t2d.zero_()
    draft.t2d[ids] = True
    draft.d2t.copy_(ids - torch.arange(DV))
    draft = draft.to(device=dev, dtype=torch.bfloat16)
    draft.freeze_embedding()
    model = OnlineEagle3Model(draft_model=draft, length=T, attention_backend=backend).to(dev)
    with tempfile.TemporaryDirectory() as wd:
        with open(os.path.join(wd, "config.json"), "w") as f:
            json.dump({"architectures": ["LlamaForCausalLM"], "model_type": "llama", "hidden_size": H, "vocab_size": V, "num_hidden_layers": 1,
                       "num_attention_heads": 4, "intermediate_size": 128}, f)
        head = TargetHead(wd)
    with torch.no_grad():
        w = torch.randn(V, H, generator=g)
        w[ids] *= 2.0                      # as bench.py: the teacher's argmax lands in the draft vocabulary for ~95 % of positions
        head.fc.weight.copy_(w)
    head.freeze_weights()
    head = head.eval().to(device=dev, dtype=torch.bfloat16)
    strategy = Eagle3TrainStrategy(model, target_head=head, ploss_decay=0.8)
    opt = BF16Optimizer(draft, lr=1e-4, max_grad_norm=0.5, total_steps=100000, warmup_ratio=0.015)
    gd = torch.Generator(device=dev).manual_seed(1000)
    t = {"input_ids": torch.randint(0, V, (B, S), device=dev, generator=gd), "attention_mask": torch.ones(B, S, dtype=torch.long, device=dev),
         "loss_mask": torch.ones(B, S, dtype=torch.long, device=dev), "hidden_state": torch.randn(B, S, 3 * H, device=dev, generator=gd).bfloat16(),
         "target": torch.randn(B, S, H, device=dev, generator=gd).bfloat16()}
    t["loss_mask"][:, -1] = 0
    tb = TrainBatch(sample_ids=[str(i) for i in range(B)], strategy="eagle3", tensors=t, metadata={"target_repr": "hidden_state"})

    def step():
        out = strategy.forward_loss(tb)
        out.loss.backward()
        opt.step()
        return out.loss

    t0 = time.time()
    for _ in range(warmup):
        loss = step()
    torch.cuda.synchronize()
    first = time.time() - t0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"impl": "reference-on-gpu", "backend": backend, "batch": B, "seq_len": S, "ttt_length": T, "ms_per_step": ms,
            "samples_per_s": B / (ms / 1e3), "loss": float(loss), "warmup_s": first, "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
            "note": "unmodified reference modules from baseline/_ref on cuda:0, CUDA-event timing"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", nargs="+", default=["flex_attention", "sdpa"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    if not os.path.isdir(os.path.join(REF, "specforge")):
        raise SystemExit("baseline/_ref missing: run __graft_entry__.build() in the build container")
    for b in a.backend:
        try:
            print(json.dumps(run(b, a.batch, a.steps, a.warmup)), flush=True)
        except Exception as exc:   # a backend that cannot run on this box is part of the answer
            import traceback
            print(json.dumps({"impl": "reference-on-gpu", "backend": b, "batch": a.batch, "error": f"{type(exc).__name__}: {str(exc)[:400]}",
                              "trace_tail": traceback.format_exc()[-600:]}), flush=True)
        import gc
        import torch
        gc.collect()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
