import time, numpy as np, torch
xs = np.random.rand(65536, 3)
pin = torch.empty((65536, 3), dtype=torch.float64, pin_memory=True); pin_np = pin.numpy()
src = torch.from_numpy(xs)
def bench(f, n=300):
    for _ in range(20): f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e6
print('torch threads', torch.get_num_threads())
print('np.copyto            %.1f us' % bench(lambda: np.copyto(pin_np, xs)))
print('torch copy_          %.1f us' % bench(lambda: pin.copy_(src)))
import ctypes
print('ctypes.memmove       %.1f us' % bench(lambda: ctypes.memmove(pin_np.ctypes.data, xs.ctypes.data, xs.nbytes)))
page = np.empty_like(xs)
print('np.copyto pageable->pageable %.1f us' % bench(lambda: np.copyto(page, xs)))
