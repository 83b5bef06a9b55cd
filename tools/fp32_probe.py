"""FP32 peak probes on a GPU box: our FFMA probe vs cuBLAS SGEMM (TF32 off)."""
import sys, torch
sys.path.insert(0, ".")
from groopm_b200 import engine
print("gmd probe TFLOP/s", engine.fp32_peak_tflops(iters=512))
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
for n in (4096, 8192):
    a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
    for _ in range(3): c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"cuBLAS sgemm {n}: {2*n**3/best/1e9:.1f} TFLOP/s ({best:.2f} ms)")
