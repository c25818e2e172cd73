# cuBLAS DGEMM rate on this GPU (the "checked baseline" / measured FP64 peak denominator).
import json, torch
dev = torch.device('cuda:0')
res = {}
for n in (2048, 4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    for _ in range(3): c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[n] = 2.0 * n ** 3 / best / 1e9
    print(f"cuBLAS DGEMM n={n}: {res[n]:.2f} TFLOP/s ({best:.3f} ms)")
# sustained
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(40): c = a @ b
e1.record(); torch.cuda.synchronize()
res['sustained_8192'] = 40 * 2.0 * n ** 3 / e0.elapsed_time(e1) / 1e9
print(f"cuBLAS DGEMM sustained n=8192 x40: {res['sustained_8192']:.2f} TFLOP/s")
json.dump(res, open('gpurun_out/dgemm_cublas.json', 'w'))
