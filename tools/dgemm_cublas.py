"""cuBLAS DGEMM ceiling on this GPU (library baseline; what the reference's ACC path calls,
reference Blas.cxx:62-75) for atrip's GEMM shapes and for a large square."""
import torch, time
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
def bench(m, n, k, batch=1, reps=10):
    a = torch.randn(batch, m, k, device=dev, dtype=torch.float64)
    b = torch.randn(batch, k, n, device=dev, dtype=torch.float64)
    for _ in range(3): torch.bmm(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): torch.bmm(a, b)
    e1.record(); torch.cuda.synchronize()
    s = e0.elapsed_time(e1) * 1e-3 / reps
    print(f"dgemm batch {batch:4d} M {m:6d} N {n:5d} K {k:5d}: {2.0*batch*m*n*k/s/1e12:7.2f} TFLOP/s  {s*1e6:9.1f} us", flush=True)
bench(8192, 8192, 8192, reps=3)
bench(4096, 4096, 4096)
for no, nv in [(40, 400), (64, 640), (100, 1000)]:
    bench(no * no, no, nv)            # one particle GEMM (reference shape)
    bench(no * no, no, nv, batch=96)  # 96 of them batched
    bench(no * no, no * 64, nv)       # N-stacked over 64 tuples
