"""cuBLAS FP64 GEMM roofline denominators (DGEMM / ZGEMM) measured with torch.matmul + CUDA events.
Writes one JSON line; recorded under profiles/ as the "of measured" FP64 peak (MEASURED_PEAKS.json has no FP64 figure)."""
import json, time, torch

def bench(dtype, n, reps=10, sustain_s=3.0):
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    flops = (8.0 if dtype.is_complex else 2.0) * n ** 3
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); k = 0
    e0.record()
    while time.time() - t0 < sustain_s:
        for _ in range(4):
            a @ b; k += 1
        torch.cuda.synchronize()
    e1.record(); e1.synchronize()
    sus = e0.elapsed_time(e1) / k
    return {"burst_tflops": flops / best * 1e-9, "sustained_tflops": flops / sus * 1e-9, "n": n}

if __name__ == "__main__":
    out = {"gpu": torch.cuda.get_device_name(0),
           "dgemm": bench(torch.float64, 8192),
           "zgemm": bench(torch.complex128, 4096)}
    print(json.dumps(out))
