"""cuBLAS DGEMM on the shapes of the delayed trailing updates (development probe; not a bench number)."""
import torch
dev = torch.device("cuda:0")
def t(m, n, k, reps=5):
    A = torch.randn(k, m, device=dev, dtype=torch.float64).t()      # column-major m x k
    B = torch.randn(n, k, device=dev, dtype=torch.float64).t()      # column-major k x n  (k contiguous)
    C = torch.randn(n, m, device=dev, dtype=torch.float64).t()
    for _ in range(2): C.addmm_(A, B, alpha=-1.0)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): C.addmm_(A, B, alpha=-1.0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"m={m} n={n} k={k}: {ms*1e3:8.1f} us  {2.0*m*n*k/ms/1e9:6.2f} TFLOP/s", flush=True)
for (m, n, k) in [(8192, 8192, 8192), (9000, 9000, 456), (9000, 9000, 912), (9000, 9000, 228), (13824, 13824, 456), (4608, 4608, 456), (9000, 456, 57), (9000, 456, 456), (2304, 2304, 456)]:
    t(m, n, k)
