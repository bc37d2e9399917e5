"""The bespoke grid-resident 3-D FFT against cuFFT (cuFFT is the bar, CudaKernels.cpp:826-829, 1228, 1255).

  ours  : k_fft_slab_fwd + k_fft_x_conv (forward x, influence function, inverse x) + k_fft_slab_inv, i.e. R2C + convolution + C2R,
          input = the int64 fixed-point charge grid, timed with CUDA events by b200md_time_phase (30 launches, warm)
  cuFFT : torch.fft.rfftn + torch.fft.irfftn on an fp32 grid of the same shape (cufftExecR2C + cufftExecC2R; the reference adds a
          separate reciprocalConvolution kernel in between, NOT included here), timed the same way
Writes gpurun_out/r02_fft_vs_cufft.md."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from openmm_b200 import systems, Engine

rows = []
for n, side in ((56, 20), (88, 28), (128, 40)):
    d = systems.water_box(side, cutoff=0.9).rounded()
    d.pme_alpha, d.pme_grid = d.pme_parameters()[0], (n, n, n)
    eng = Engine(d)
    eng.compute()
    ours = eng.time_phase("pme_fft_conv", 30)*1e3
    eng.close()
    x = torch.randn(n, n, n, device="cuda")
    for _ in range(5):
        y = torch.fft.irfftn(torch.fft.rfftn(x), s=(n, n, n))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        y = torch.fft.irfftn(torch.fft.rfftn(x), s=(n, n, n))
    e1.record(); torch.cuda.synchronize()
    cufft = e0.elapsed_time(e1)/30*1e3
    gbytes = (12*n**3 + 48*n*n*(n//2 + 1))/1e9
    rows.append((n, d.natoms, ours, cufft, gbytes/(ours*1e-6)))
    print(n, ours, cufft, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/r02_fft_vs_cufft.md", "w") as f:
    f.write("# bespoke FFT (R2C + convolution + C2R, 3 launches) vs cuFFT (R2C + C2R through torch.fft, no convolution), B200, CUDA events, 30 warm launches\n\n")
    f.write("| grid | ours (us) | cuFFT R2C+C2R (us) | ours / cuFFT | ours: algorithmic GB/s (12 G + 48 H bytes) |\n|---|---|---|---|---|\n")
    for n, na, o, c, bw in rows:
        f.write("| %d^3 | %.1f | %.1f | %.2f | %.0f |\n" % (n, o, c, o/c, bw))
print(open("gpurun_out/r02_fft_vs_cufft.md").read())
