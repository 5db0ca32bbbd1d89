#!/usr/bin/env python3
"""Developer tool: per-CTA cycle counts of the last weight-gradient launch of one training step (needs a library
built with `touch nonrigid_nerf_b200/csrc/wgrad.cu; make -C nonrigid_nerf_b200/csrc EXTRA=-DNRN_TRACE`)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle.nrnerf_oracle as O  # noqa: E402  (developer script, not a product path)
from tests import helpers  # noqa: E402
from nonrigid_nerf_b200 import train as T, _lib  # noqa: E402


def main():
    dev = "cuda:0"
    seed, n = 5, 1024
    coarse, fine, bender, _ = helpers.build_models(O, seed, dev)
    r = O.make_rays(seed, n)
    kw = dict(network_query_fn=None, perturb=1.0, N_importance=128, network_fine=fine, N_samples=64, network_fn=coarse, ray_bender=bender,
              use_viewdirs=False, white_bkgd=False, raw_noise_std=1.0, ndc=False, lindisp=False)
    tgt = r["target"].to(dev)
    for it in range(3):
        rgb, _, _, ex = T.render(r["rays_o"].to(dev), r["rays_d"].to(dev), chunk=32768, near=r["near"], far=r["far"],
                                 additional_pixel_information={"ray_bending_latents": r["latents"].to(dev)}, detailed_output=True, **kw)
        (((rgb - tgt) ** 2).mean() + ((ex["rgb0"] - tgt) ** 2).mean()).backward()
    torch.cuda.synchronize()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    buf = (ctypes.c_longlong * (192 * 4))()
    lib.dbg_wgrad_prof_read(buf)
    rows = [(buf[4 * i], buf[4 * i + 1], buf[4 * i + 2], buf[4 * i + 3], i) for i in range(192) if buf[4 * i + 3]]
    print("last wgrad launch (the coarse pass: 512 tiles): job, CTAs, tiles/CTA, kcycles until MMAs done (min..max), total (max)")
    for job in sorted(set(x[0] for x in rows)):
        sel = [x for x in rows if x[0] == job]
        print(f"job {job:2d}: {len(sel):3d} CTAs  tiles {min(x[1] for x in sel):4d}..{max(x[1] for x in sel):4d}  "
              f"mma-done {min(x[2] for x in sel) / 1e3:7.1f}..{max(x[2] for x in sel) / 1e3:7.1f}  total {max(x[3] for x in sel) / 1e3:7.1f}")


if __name__ == "__main__":
    main()
