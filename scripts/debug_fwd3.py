"""Developer tool: locate where field_fwd3.cu's training-mode output departs from field_fwd.cu's (per tile / per stash image)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle.nrnerf_oracle as O
from tests import helpers
from nonrigid_nerf_b200 import autograd as ag, ops, _lib

dev = torch.device("cuda:0")
lib = _lib.load()
coarse, fine, bender, _ = helpers.build_models(O, 19, dev, True)
names = [("E", 0, 16384)] + [(f"H{l+1}", 16384 + l * 65536, 65536) for l in range(8)] + [("Bin", 16384 + 8 * 65536, 6 * 2048)]
for n, s, with_b in ((2500, 64, False), (2500, 64, True), (5000, 128, False)):
    r = O.make_rays(19, n)
    rays = helpers.rays8(r, dev)
    z = ops.sample_coarse(rays, s, None, False)
    coarse.ray_bender = (bender if with_b else None,)
    lat = r["latents"].to(dev) if with_b else None
    for rep in range(3):
        got = {}
        for kind in (1, 3):
            _lib.check(lib.nrn_select_forward_kernel(kind), "select")
            stash = torch.zeros(lib.nrn_stash_bytes(n, s), dtype=torch.uint8, device=dev)
            raw_t, _ = ops.field_forward(rays, z, lat, ops.pack_nerf(coarse), ops.pack_bender(bender) if with_b else None, 5, None, None, None, True, stash)
            _lib.device_error_check()
            got[kind] = (raw_t, stash)
        a, b = got[1], got[3]
        n_tiles = (n * s + 127) // 128
        bad_rows = (a[0].reshape(-1, 5) != b[0].reshape(-1, 5)).any(-1)
        tiles_bad = torch.unique(torch.nonzero(bad_rows).flatten() // 128)
        print(f"n={n} s={s} bender={with_b} rep={rep}: raw rows differing {int(bad_rows.sum())} / {bad_rows.numel()}; tiles {tiles_bad.tolist()[:12]} (pairs {sorted(set((tiles_bad // 2).tolist()))[:12]}) of {n_tiles}")
        sa = a[1][: n_tiles * 634880].reshape(n_tiles, 634880)
        sb = b[1][: n_tiles * 634880].reshape(n_tiles, 634880)
        for nm, off, ln in names:
            d = (sa[:, off:off + ln] != sb[:, off:off + ln])
            if bool(d.any()):
                t = torch.nonzero(d.any(-1)).flatten()
                first = d[t[0]].nonzero().flatten()
                chunks = torch.unique(first // 2048)
                print(f"   stash {nm}: {int(d.sum())} bytes differ in tiles {t.tolist()[:8]}; first tile: chunks {chunks.tolist()[:20]} rows {torch.unique((first % 2048) // 16).tolist()[:8]}")
lib.nrn_select_forward_kernel(1)
