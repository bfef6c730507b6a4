"""First timing of the DFlash step (first correct CUDA version: CUDA-core block attention, not tcgen05 yet).
python tools/dflash_bench.py [B] [S] [N]      defaults = BASELINE config 4 per GPU: 4 x 2048 context tokens, 512 anchors x 16."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from specforge_b200.dflash import DFlashDims, DFlashEngine, sample_anchor_positions   # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    N = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    dims = DFlashDims(hidden_size=4096, intermediate_size=12288, num_heads=32, num_kv_heads=8, head_dim=128, num_layers=5,
                      num_target_feats=5, vocab_size=151936, block_size=16, mask_token_id=151669, rope_theta=1e6)
    dev = torch.device("cuda", 0)
    eng = DFlashEngine(dims, batch=B, seq_len=S, num_blocks=N, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    eng.params.copy_((torch.randn(eng.n_params, device=dev, generator=g) * 0.02).bfloat16())
    for n in eng.names:
        if n.endswith("norm.weight") or n.endswith("layernorm.weight"):
            eng.param_view(n).fill_(1.0)
    eng.set_frozen(embed_tokens=torch.randn(dims.vocab_size, 4096, device=dev, generator=g) * 0.02,
                   lm_head=torch.randn(dims.vocab_size, 4096, device=dev, generator=g) * 0.02)
    batch = {"input_ids": torch.randint(0, 151000, (B, S), device=dev, generator=g),
             "hidden_states": torch.randn(B, S, 5 * 4096, device=dev, generator=g).bfloat16(), "loss_mask": torch.ones(B, S, device=dev)}
    anchors, keep = sample_anchor_positions(batch["loss_mask"], N)
    Mq, Mc, H, I, V, A, KV, L = B * anchors.shape[1] * 16, B * S, 4096, 12288, 151936, 4096, 1024, 5
    fwd = 2 * Mc * 5 * H * H + L * (2 * Mq * H * (A + 2 * KV) + 2 * Mc * H * 2 * KV + 2 * Mq * A * H + 6 * Mq * H * I) + 2 * Mq * H * V
    attn = L * 4 * B * 32 * 128 * 16 * float((anchors.float() + 16).sum() / B)
    flops = 3 * fwd - 2 * Mq * H * V + 3.5 * attn       # no weight gradient for the frozen head
    t0 = time.time()
    eng.forward(batch, anchors, keep); eng.backward()
    torch.cuda.synchronize()
    first = time.time() - t0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 2
    e0.record()
    for _ in range(steps):
        eng.forward(batch, anchors, keep); eng.backward()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"workload": f"DFlash draft step, Qwen3-8B dims, B={B} S={S} anchors={anchors.shape[1]} block=16 (fwd+bwd, no optimizer)",
                      "ms_per_step": ms, "first_step_s": first, "samples_per_s": B / (ms / 1e3), "algorithmic_tflop_per_step": flops / 1e12,
                      "tflops": flops / 1e12 / (ms / 1e3), "loss": float(eng.loss), "workspace_gb": eng.workspace_bytes / 1e9,
                      "note": "first correct version: block attention on CUDA cores"}), flush=True)


if __name__ == "__main__":
    main()
