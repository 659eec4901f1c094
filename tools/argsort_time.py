import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import pints_b200 as pb
for n in (4096, 16384, 65536, 262144):
    s = pb.DeviceArgsort(n, None)
    v = torch.randn(n, dtype=torch.float64, device=s.device)
    def T(f, reps=200):
        for _ in range(10): f()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): f()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e6
    print(n, 'native %.1f us' % T(lambda: s(v)), 'torch %.1f us' % T(lambda: torch.sort(v)))
