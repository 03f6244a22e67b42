"""Write-only / read-only / copy bandwidth of the box's HBM with library kernels (context for the HBM-bound conv rows:
MEASURED_PEAKS.json's hbm_gbs is a copy figure, i.e. read + write)."""
import torch
n = 1 << 30                      # 4 GiB of float32: far beyond the 126 MB L2
x = torch.empty(n, dtype=torch.float32, device="cuda"); y = torch.empty_like(x)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts) * 1e-3
tw = t(lambda: x.zero_()); print("write-only (memset)   %.0f GB/s" % (4 * n / tw / 1e9))
tf = t(lambda: x.fill_(1.5)); print("write-only (fill kernel) %.0f GB/s" % (4 * n / tf / 1e9))
tr = t(lambda: x.sum()); print("read-only (sum)       %.0f GB/s" % (4 * n / tr / 1e9))
tc = t(lambda: y.copy_(x)); print("copy (read + write)   %.0f GB/s" % (8 * n / tc / 1e9))
