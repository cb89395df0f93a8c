import sys, os
sys.path[:0] = ['.', './gauss-process_b200']
import numpy as np, torch
from b200id import fir
from b200id.runtime import require_cuda
dev = require_cuda()
n = 10_000_000
gen = torch.Generator(device=dev); gen.manual_seed(0)
u = torch.randn(n, dtype=torch.float64, device=dev, generator=gen)
y = torch.randn(n, dtype=torch.float64, device=dev, generator=gen)
g = torch.randn(1024, dtype=torch.float64, device=dev, generator=gen) * 0.01
for _ in range(3): fir.fir_eval_device(u, y, g, 10, 0.0)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20): sums, _ = fir.fir_eval_device(u, y, g, 10, 0.0)
e.record(); torch.cuda.synchronize()
print(os.environ.get("B200ID_LIB", "default"), "fir_eval 1e7 x 1024 taps: %.4f ms" % (s.elapsed_time(e) / 20), sums.cpu().numpy())
