"""Host-side cost of one bench step (no GPU sync inside the step except the reference-mandated one):
wall time per step with and without a trailing synchronize, plus a cProfile of 20 steps."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gscodec_studio_b200 as gs  # noqa: E402
from gscodec_studio_b200.compression_simulation import CompressionSimulation  # noqa: E402
from gscodec_studio_b200.losses import photometric_loss  # noqa: E402

dev = torch.device("cuda:0")
import os as _os
raw, viewmat, K, target, n_total = bench.build_workload(0, 1, int(_os.environ.get("SCENE_GRID", "3")))
params = {k: v.to(dev).requires_grad_(True) for k, v in raw.items()}
viewmat, K, target = viewmat.to(dev), K.to(dev), target.to(dev)
sim = CompressionSimulation(entropy_model_enable=False, device=dev)


def step():
    for p in params.values():
        p.grad = None
    new, bits = sim.simulate_compression(params, 100, fuse_activations=True)
    colors = (new["sh0"], new["shN"])
    rc, ra, meta = gs.rasterization(new["means"], new["quats"], new["scales_act"], new["opacities_act"], colors, viewmat, K,
                                    bench.W_IMG, bench.H_IMG, sh_degree=3, packed=False, near_plane=0.01, far_plane=100.0,
                                    radius_clip=3.0)
    loss = photometric_loss(rc, target, ssim_lambda=0.0)
    loss.backward()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"async: host {1e3 * (t1 - t0) / 20:.3f} ms/step, with final sync {1e3 * (t2 - t0) / 20:.3f} ms/step")
t0 = time.perf_counter()
for _ in range(20):
    step().item()
t1 = time.perf_counter()
print(f"sync every step: {1e3 * (t1 - t0) / 20:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step().item()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
pstats.Stats(pr).sort_stats("tottime").print_stats(40)
