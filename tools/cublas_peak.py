"""The GEMM behind MEASURED_PEAKS.json's bf16 figure (torch.matmul 8192^3, cuBLAS), for ONE `ncu --set full` capture that
calibrates sm__pipe_tensor_cycles_active: what the metric reads on a kernel that runs at the measured peak."""
import torch

a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    c = a @ b
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(10):
    c = a @ b
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 10
print("matmul 8192^3 bf16: %.3f ms = %.1f TFLOP/s" % (ms, 2 * 8192 ** 3 / ms / 1e9))
