"""TF32 dense tensor peak measured the way MEASURED_PEAKS.json measures bf16: torch.matmul 8192^3 with TF32 allowed, best of 10
(burst) and back to back for 4 s (sustained).  Library GEMM as the yardstick, not part of the product path.
usage: python profiles/measure_tf32_peak.py > gpurun_out/tf32_peak.json"""
import json, time
import torch
torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
for _ in range(3):
    a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
t0 = time.time(); k = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(20):
        a @ b
    k += 20
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
fl = 2.0 * n ** 3
print(json.dumps({"tf32_tflops": fl / best / 1e9, "tf32_tflops_sustained": fl * k / e0.elapsed_time(e1) / 1e9,
                  "how": "torch.matmul fp32 with allow_tf32, 8192^3, best of 10 / 4 s back to back", "gpu": torch.cuda.get_device_name(0)}))
