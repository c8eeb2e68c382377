"""TEST INFRASTRUCTURE — recipe that makes the reference's hot-path sources available where /root/reference is absent.

The reference is a pure-Python repo without package metadata (requirements.txt = matplotlib only), so there is nothing
to ``pip install``; the GPU box has no /root/reference. This script copies, UNMODIFIED, the few reference files the
per-sample path imports into the git-ignored (but gpurun-shipped) directory ``oracle/_ref/``; ``oracle/reference_loader``
looks there when /root/reference does not exist. ``__graft_entry__.build()`` runs it in the build container.
Nothing under oracle/_ref/ is tracked, and nothing in the product imports it: it serves ``bench.py --impl reference`` /
``cpu_baseline`` (kind "reference") and nothing else.

    python -m oracle.fetch_ref
"""
from __future__ import annotations

import os
import shutil

SRC = os.environ.get("UA_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = [
    "Uni_Adapter.py", "dota.py", "dota_mixture.py",
    "utils/__init__.py", "utils/utils.py", "utils/math_utils.py",
    "visualize/visualization.py",
    "models/point_encoder.py",
    "models/ulip/pointbert/misc.py", "models/ulip/pointbert/dvae.py", "models/ulip/pointbert/point_encoder.py",
    "models/openshape/pointnet_util.py", "models/openshape/ppta.py",
    "LICENSE",
]


def fetch() -> bool:
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    for rel in FILES:
        src = os.path.join(SRC, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    return True


if __name__ == "__main__":
    print("oracle/_ref ready" if fetch() else "no reference sources found")
