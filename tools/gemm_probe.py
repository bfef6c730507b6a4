"""GPU probe: checks every tcgen05 GEMM variant against torch.matmul and times it.
Each variant runs in its own subprocess under a timeout so a trapped/hung kernel cannot take the box down."""
import json, os, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(G, am, bm, epi, M, N, K, timing):
    import torch
    from specforge_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    a = A if am == 0 else A.t().contiguous()
    b = B if bm == 0 else B.t().contiguous()
    R = torch.randn(M, N, device=dev).bfloat16() if epi == 1 else None
    out0 = torch.randn(M, N, device=dev, dtype=torch.float32) if epi == 3 else None
    ref = A.float() @ B.float().t()
    if epi == 1: ref = ref.bfloat16().float() + R.float()
    if epi == 3: ref = ref + out0
    out = out0.clone() if epi == 3 else None
    out = ops.gemm(a, b, a_major=am, b_major=bm, out=out, residual=R, epi=epi, cta_group=G)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    res = {"G": G, "am": am, "bm": bm, "epi": epi, "M": M, "N": N, "K": K, "max_err": err, "ref_max": scale,
           "ok": bool(err <= 2e-2 * scale + 1e-3)}
    if timing:
        it = 10
        outs = ops.gemm(a, b, a_major=am, b_major=bm, epi=0 if epi in (0, 1) else 2, cta_group=G)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        for _ in range(it):
            ops.gemm(a, b, a_major=am, b_major=bm, out=outs, epi=0 if epi in (0, 1) else 2, cta_group=G)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / it
        res["ms"] = ms; res["tflops"] = 2.0 * M * N * K / ms / 1e9
        s.record()
        for _ in range(it):
            torch.matmul(A, B.t())
        e.record(); torch.cuda.synchronize()
        res["cublas_tflops"] = 2.0 * M * N * K / (s.elapsed_time(e) / it) / 1e9
    print("RESULT " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        G, am, bm, epi, M, N, K, timing = map(int, sys.argv[2:10])
        one(G, am, bm, epi, M, N, K, timing)
        sys.exit(0)
    cases = []
    for G in (1, 2):
        for (am, bm) in ((0, 0), (0, 1), (1, 1)):
            cases.append((G, am, bm, 0, 512, 512, 256, 0))       # small, exact tiles
            cases.append((G, am, bm, 2, 384, 320, 200, 0))       # ragged M/N/K, fp32 out
            cases.append((G, am, bm, 0, 8192, 4096, 4096, 1))    # perf
    cases += [(2, 0, 0, 1, 1024, 768, 512, 0), (2, 1, 1, 3, 1024, 768, 4096, 0), (1, 0, 0, 1, 300, 264, 64, 0),
              (2, 0, 0, 0, 16384, 32000, 4096, 1), (2, 1, 1, 2, 4096, 12288, 16384, 1), (2, 0, 1, 0, 16384, 4096, 12288, 1)]
    results = []
    for c in cases:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, "one"] + [str(x) for x in c], capture_output=True, text=True,
                               timeout=150)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            if line:
                r = json.loads(line[0][7:])
            else:
                r = {"case": c, "error": (p.stdout[-800:] + p.stderr[-1500:])}
        except subprocess.TimeoutExpired:
            r = {"case": c, "error": "timeout"}
        r["wall_s"] = round(time.time() - t0, 1)
        print(json.dumps(r), flush=True)
        results.append(r)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gemm_probe.json"), "w") as f:
        json.dump(results, f, indent=1)
