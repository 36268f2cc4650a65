#!/usr/bin/env python
"""Vendor the seven reference files that construct netG into baseline/_ref/ (git-ignored; travels to the GPU box with
the repo snapshot like the built .so does) so that `bench.py --impl reference` can time the UNMODIFIED reference
module -- `define_G` of Module2/models/networks.py:123-201 -- on the box's host cores.

    python tools/prep_ref.py            # copies from /root/reference when it exists; no-op otherwise

The files are copied byte for byte and never committed.  Run by `__graft_entry__.build()` in the build container.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("AP_REFERENCE_ROOT", "/root/reference")
DEST = os.path.join(ROOT, "baseline", "_ref")
FILES = [
    "Module2/models/__init__.py", "Module2/models/base_model.py", "Module2/models/networks.py", "Module2/models/facenet.py",
    "Module2/intrinsic_flow_models/__init__.py", "Module2/intrinsic_flow_models/networks.py",
    "Module2/intrinsic_flow_models/modules.py",
]


def prep(verbose: bool = True) -> bool:
    if not os.path.isdir(REFERENCE):
        if verbose:
            print(f"prep_ref: {REFERENCE} not present; baseline/_ref left as it is")
        return os.path.exists(os.path.join(DEST, FILES[2]))
    for rel in FILES:
        src, dst = os.path.join(REFERENCE, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    if verbose:
        print(f"prep_ref: {len(FILES)} reference files -> {DEST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if prep() else 1)
