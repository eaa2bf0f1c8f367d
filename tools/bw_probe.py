"""Reference points for the HBM roofline on this box: pure write (memset), pure read (sum), copy; CUDA events, large buffers."""
import torch

dev = "cuda"
n = 2 * 1024 ** 3            # 2 GiB
a = torch.empty(n, dtype=torch.uint8, device=dev)
b = torch.empty(n, dtype=torch.uint8, device=dev)
af = a.view(torch.float32)


def t(fn, reps=5):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e-3


for _ in range(2):
    a.zero_(); b.copy_(a)
tz = t(lambda: a.zero_())
tc = t(lambda: b.copy_(a))
tr = t(lambda: af.sum())
print(f"memset  {n / tz / 1e9:7.0f} GB/s written")
print(f"copy    {2 * n / tc / 1e9:7.0f} GB/s read+written")
print(f"reduce  {n / tr / 1e9:7.0f} GB/s read")
