"""Write-only vs copy HBM bandwidth on this GPU (context for the roofline of a write-dominated kernel)."""
import torch, json
n = 1 << 28  # 1 GiB of f32
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
res = {}
t = timeit(lambda: a.fill_(1.5)); res["fill_gbs"] = 4 * n / t / 1e6
t = timeit(lambda: a.zero_()); res["memset_gbs"] = 4 * n / t / 1e6
t = timeit(lambda: b.copy_(a)); res["copy_gbs_rw"] = 8 * n / t / 1e6
t = timeit(lambda: torch.add(a, 1.0, out=b)); res["add_gbs_rw"] = 8 * n / t / 1e6
c = torch.empty(n // 4, dtype=torch.float32, device="cuda")
t = timeit(lambda: c.fill_(2.0)); res["fill_256MB_gbs"] = n / t / 1e6
print(json.dumps(res))
