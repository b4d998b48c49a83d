"""Quick timing of the cfg2 forward / backward kernels (diagnostic)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shwd
dev = torch.device("cuda:0")
B, N, L = 32, 1024, 100
KIND = sys.argv[1] if len(sys.argv) > 1 else "geodesic"
P = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
torch.manual_seed(1234)
x = torch.nn.functional.normalize(torch.randn(B, N, 3), dim=-1).to(dev).requires_grad_(True)
y = torch.nn.functional.normalize(torch.randn(B, N, 3), dim=-1).to(dev).requires_grad_(True)
fs, bs = [], []
for it in range(12):
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    res = shwd.entropic_ot(x, y, KIND, P, 0.01, L, center=(KIND == "geodesic"))
    e1.record()
    res.cost.sum().backward()
    e2.record()
    torch.cuda.synchronize()
    if it >= 2:
        fs.append(e0.elapsed_time(e1)); bs.append(e1.elapsed_time(e2))
import statistics, subprocess
f, b = min(fs), min(bs)
try:
    clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
except OSError:
    clk = "?"
print(KIND, P, "fwd %.2f ms  bwd %.2f ms  -> %.0f pairs/s   (median fwd %.2f bwd %.2f; idle clock/power %s)  status %d" % (
    f, b, B / ((f + b) * 1e-3), statistics.median(fs), statistics.median(bs), clk, res.status()))
