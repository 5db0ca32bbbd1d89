#!/usr/bin/env python3
"""Developer timing of the fused field kernel alone (CUDA events, L2 flushed between iterations)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle.nrnerf_oracle as O  # noqa: E402  (developer script, not a product path)
from tests import helpers  # noqa: E402
from nonrigid_nerf_b200 import autograd as ag, ops, _lib  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    coarse, fine, bender, _ = helpers.build_models(O, 1, dev, True)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for n, s in ((1024, 64), (1024, 128), (8192, 128), (65536, 128)):
        r = O.make_rays(1, n)
        rays = helpers.rays8(r, dev)
        lat = r["latents"].to(dev)
        z = ops.sample_coarse(rays, s, None, False)
        for with_b in (True, False):
            net = coarse
            net.ray_bender = (bender if with_b else None,)
            for _ in range(3):
                ag.field_rays(net, rays, z, lat if with_b else None, False)
            _lib.device_error_check()
            ts = []
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ag.field_rays(net, rays, z, lat if with_b else None, False)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            flop = n * s * (O.FLOP_PER_POINT if with_b else 984576)
            print(f"n={n:6d} S={s:3d} bender={int(with_b)}: {ms:8.3f} ms  {n * s / ms / 1e3:9.1f} Mpts/s  {flop / ms / 1e9:8.1f} TFLOP/s")
        coarse.ray_bender = (bender,)


if __name__ == "__main__":
    main()
