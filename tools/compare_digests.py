#!/usr/bin/env python
"""Cross-check of two runs of the same workload on different GPU counts (bench.py --digest): relative L2 difference of the
sampled fields (every 997th node, matched by global id).   python tools/compare_digests.py gpurun_out/s256 8 1"""
import glob
import sys

import numpy as np


def load(prefix, n):
    parts = [np.load(f) for f in sorted(glob.glob(f"{prefix}_n{n}_r*.npz"))]
    assert len(parts) == n, (prefix, n, len(parts))
    gid = np.concatenate([p["gid"] for p in parts])
    o = np.argsort(gid)
    return gid[o], {k: np.concatenate([p[k] for p in parts])[o] for k in "Vuv"}, [int(p["iters"]) for p in parts], sum(p["sums"] for p in parts)


prefix, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ga, fa, ia, sa = load(prefix, a)
gb, fb, ib, sb = load(prefix, b)
assert np.array_equal(ga, gb)
print(f"{ga.size} sampled nodes; Gummel iterations {ia[0]} vs {ib[0]}")
for k in "Vuv":
    print(f"  {k}: rel-L2 difference {np.linalg.norm(fa[k] - fb[k]) / np.linalg.norm(fb[k]):.3e}")
print("  whole-field sums / squared norms, relative difference:", np.abs(sa - sb) / np.abs(sb))
