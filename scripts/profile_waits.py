#!/usr/bin/env python3
"""Developer tool: where do the roles of the fused forward kernel wait?  (NRN_DEBUG_MODE=9 makes CTA 0 accumulate the cycles
each role spends blocked on each kind of barrier; see g_fwd1_prof in csrc/field_fwd.cu.  profiles/r02_profile_waits.log also
holds the rows of the shared-slab kernel that was built and removed in round 2)"""
import ctypes
import os
import sys

import torch

os.environ["NRN_DEBUG_MODE"] = "9"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle.nrnerf_oracle as O  # noqa: E402  (developer script, not a product path)
from tests import helpers  # noqa: E402
from nonrigid_nerf_b200 import autograd as ag, ops, _lib  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    lib = _lib.load()
    raw = ctypes.CDLL(_lib.LIB_PATH)
    coarse, fine, bender, _ = helpers.build_models(O, 1, dev, True)
    n, s = 16384, 128
    r = O.make_rays(1, n)
    rays = helpers.rays8(r, dev)
    z = ops.sample_coarse(rays, s, None, False)
    lat = r["latents"].to(dev)
    ghz = 1.965
    for with_b in (False, True):
        coarse.ray_bender = (bender if with_b else None,)
        for _ in range(2):
            ag.field_rays(coarse, rays, z, lat if with_b else None, False)
        torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 16)()
        raw.nrn_debug_profile_fwd1(buf)
        v = [int(x) for x in buf]
        pairs = max(v[11], 1)
        us = lambda c: c / ghz / 1e3 / pairs
        print(f"bender={int(with_b)}: per tile pair {us(v[0]):7.2f} us | issuer waits: a_ready {us(v[1]):6.2f}  w_full {us(v[3]):6.2f}  "
              f"(issuing/other {us(v[0] - v[1] - v[3]):6.2f}) | producer waits w_empty {us(v[4]):6.2f} | epilogue WG: total {us(v[5]):7.2f} waits d_full {us(v[6]):6.2f} "
              f"| {v[10] / pairs:.0f} slabs/pair")
    coarse.ray_bender = (bender,)



if __name__ == "__main__":
    main()
