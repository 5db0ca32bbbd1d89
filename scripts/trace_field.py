#!/usr/bin/env python3
"""Developer tool: handshake time line of the fused field kernel (needs a library built with
`touch nonrigid_nerf_b200/csrc/field_fwd2.cu; make -C nonrigid_nerf_b200/csrc EXTRA=-DNRN_TRACE`).  Prints the
globaltimer-stamped events of cluster 0's third work group: who waited for whom, and for how long."""
import ctypes
import os
import sys

import torch

os.environ.setdefault("NRN_PAIR", "1")   # the instrumented kernel is the CTA-pair forward; read once when the library loads

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle.nrnerf_oracle as O  # noqa: E402  (developer script, not a product path)
from tests import helpers  # noqa: E402
from nonrigid_nerf_b200 import autograd as ag, ops, _lib  # noqa: E402

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
    pair = os.environ.get("NRN_PAIR", "1") != "0"
    (lib.dbg_trace2_reset if pair else lib.dbg_trace_reset)()
    ag.field_rays(coarse, rays, z, None, False)
    torch.cuda.synchronize()
    if pair:
        buf = (ctypes.c_ulonglong * 8192)()
        k = lib.dbg_trace2_read(buf, 8192)
        ev = sorted((b >> 20, (b >> 19) & 1, (b >> 12) & 127, (b >> 6) & 63, (b >> 4) & 3, b & 15) for b in buf[:k] if b)
        t0 = ev[0][0]
        names = {1: "issuer: a_ready seen", 6: "issuer: turn taken", 2: "issuer: slab ready", 3: "issuer: d_full commit issued", 4: "epilogue: d_full seen", 5: "epilogue: published"}
        for t, cta, e, step, slot, j in ev:
            if cta == 0:
                print(f"{t - t0:8d} ns  cta{cta} slot{slot} step{step:2d}  {names.get(e, e)}")
        return
    raise SystemExit("only the CTA-pair forward kernel (NRN_PAIR unset or 1) is instrumented")


if __name__ == "__main__":
    main()
