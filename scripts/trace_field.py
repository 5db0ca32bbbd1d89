#!/usr/bin/env python3
"""Developer tool: handshake time line of the fused field kernel (needs a library built with
`make -C nonrigid_nerf_b200/csrc EXTRA=-DNRN_TRACE`).  Prints the globaltimer-stamped events of cluster 0's third
work group: who waited for whom, and for how long."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle.nrnerf_oracle as O  # noqa: E402  (developer script, not a product path)
from tests import helpers  # noqa: E402
from nonrigid_nerf_b200 import autograd as ag, ops, _lib  # noqa: E402

EV = {1: "issuer: a_ready seen", 8: "issuer: w_full seen", 2: "issuer: w_peer seen", 3: "issuer: d_full commit issued",
      4: "epilogue: d_full seen", 5: "epilogue: published", 6: "producer: w_empty seen", 7: "relay: w_full seen"}


def main():
    dev = torch.device("cuda:0")
    coarse, fine, bender, _ = helpers.build_models(O, 1, dev, True)
    coarse.ray_bender = (None,)
    n, s = 16384, 128
    r = O.make_rays(1, n)
    rays = helpers.rays8(r, dev)
    z = ops.sample_coarse(rays, s, None, False)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for _ in range(2):
        ag.field_rays(coarse, rays, z, None, False)
    torch.cuda.synchronize()
    lib.dbg_trace_reset()
    ag.field_rays(coarse, rays, z, None, False)
    torch.cuda.synchronize()
    prof = (ctypes.c_longlong * 8)()
    lib.dbg_prof_read(prof)
    tot = max(prof[0], 1)
    print("issuer cycles: total %d | other %.1f%% | wait a_ready %.1f%% | wait weights %.1f%% | issue MMAs %.1f%% | commit+advance %.1f%%" % (
        tot, 100 * prof[1] / tot, 100 * prof[2] / tot, 100 * prof[3] / tot, 100 * prof[4] / tot, 100 * prof[5] / tot))
    buf = (ctypes.c_ulonglong * 8192)()
    k = lib.dbg_trace_read(buf, 8192)
    ev = sorted((b >> 20, (b >> 19) & 1, (b >> 12) & 127, (b >> 6) & 63, (b >> 4) & 3, b & 15) for b in buf[:k] if b)
    k = len(ev)
    t0 = ev[0][0]
    print(f"{k} events")
    for t, cta, e, step, slot, j in ev:
        if os.environ.get("NRN_TRACE_EVENTS") and (6 <= step <= 9 or e == 5):
            print(f"{t - t0:8d} ns  cta{cta} step{step:2d} slot{slot} j{j}  {EV.get(e, e)}")


if __name__ == "__main__":
    main()
