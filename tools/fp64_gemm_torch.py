"""cuBLAS DGEMM throughput via torch.matmul (float64) -- the FP64 roofline denominator."""
import torch, time, json
torch.backends.cuda.matmul.allow_tf32 = False
out = {}
for n in (2048, 4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(3):
        c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[n] = 2.0 * n ** 3 / best * 1e-9
    print(f"DGEMM n={n}: {best:.3f} ms  {out[n]:.2f} TFLOP/s", flush=True)
print(json.dumps(out))
