#!/usr/bin/env python3
"""Recipe for oracle/_ref/: the UNMODIFIED reference sources of the hot path, copied byte for byte from the read-only
checkout so that they travel to the GPU box (which has no /root/reference) as the CPU baseline of bench.py.

    python oracle/make_ref.py            # run in the build container; __graft_entry__.build() calls it

oracle/_ref/ is git-ignored (reference sources never enter this repository's history) but not gpurun-ignored.  Nothing
under nonrigid_nerf_b200/ imports it: it is test / measurement infrastructure, like the rest of oracle/.
"""
import hashlib
import os
import shutil
import sys

REF = "/root/reference"
FILES = ("train.py", "run_nerf_helpers.py")


def main() -> int:
    here = os.path.dirname(os.path.abspath(__file__))
    dst = os.path.join(here, "_ref")
    if not os.path.isdir(REF):
        print(f"make_ref: {REF} not present (GPU box?): keeping whatever {dst} holds")
        return 0
    os.makedirs(dst, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(REF, f), os.path.join(dst, f))
    with open(os.path.join(dst, "SHA256SUMS"), "w") as out:
        for f in FILES:
            out.write(f"{hashlib.sha256(open(os.path.join(dst, f), 'rb').read()).hexdigest()}  {f}\n")
    print(f"make_ref: copied {', '.join(FILES)} -> {dst}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
