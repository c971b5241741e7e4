"""Measure cuBLAS GEMM peaks (fp64 / fp32 / tf32) on the box: roofline denominators that
MEASURED_PEAKS.json does not carry. Library calls, never on the product path."""
import json, sys, torch

def gemm_tflops(dtype, n, tf32=False, reps=5):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    for _ in range(2):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / best * 1e-9

out = {"gpu": torch.cuda.get_device_name(0)}
out["fp64_tflops"] = gemm_tflops(torch.float64, 8192)
out["fp32_tflops"] = gemm_tflops(torch.float32, 8192, tf32=False)
out["tf32_tflops"] = gemm_tflops(torch.float32, 8192, tf32=True)
# batched potrf via cuSOLVER/MAGMA for context: 50k x 128x128 f32
k = torch.randn(4096, 128, 128, device="cuda")
k = k @ k.transpose(1, 2) + 128 * torch.eye(128, device="cuda")
torch.linalg.cholesky(k); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); torch.linalg.cholesky(k); e1.record(); torch.cuda.synchronize()
out["torch_cholesky_4096x128_f32_ms"] = e0.elapsed_time(e1)
print(json.dumps(out))
