"""Stress the batched edit for nondeterminism: the same edit many times on fresh pools; on a canonical mismatch walk
both DAGs top-down and report the first differing node."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vkhashdag_b200 as v
from oracle import bindings as B
from vkhashdag_b200 import abi

O = B.Oracle()
cfg = abi.default_config(level_count=10, top_level_count=9)
edits = [abi.sphere((512, 512, 512), 341 ** 2)]
op = O.pool(cfg)
oroot = op.edit_batch(abi.NULL, edits)
exp = op.canonical(oroot)
NL = cfg.node_levels


def node(words, ptr, leaf):
    if leaf:
        return [int(words[ptr]), int(words[ptr + 1])]
    m = int(words[ptr]) & 0xFF
    return [m] + [int(words[ptr + 1 + i]) for i in range(bin(m).count("1"))]


def first_diff(gw, gp, ow, op_, level, path):
    """memoised subtree equality is overkill here: walk until the first structural difference."""
    leaf = level == NL - 1
    a, b = node(gw, gp, leaf), node(ow, op_, leaf)
    if leaf or a[0] != b[0]:
        return None if a == b else (level, path, a, b)
    for i, (ca, cb) in enumerate(zip(a[1:], b[1:])):
        d = first_diff(gw, ca, ow, cb, level + 1, path + [i])
        if d:
            return d
    return None


bad = 0
N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
oview = op.words_np()
for it in range(N):
    dev = v.DAGNodePool(cfg)
    root = dev.EditBatch(abi.NULL, edits)
    st = dev.last_stats
    m = O.pool(cfg)
    ranges, bw = dev.Download()
    for off, words in ranges.items():
        m.words_np(off, len(words))[:] = words
    got = O.canonical(m.words_ptr, NL, root)
    if got != exp:
        bad += 1
        sys.setrecursionlimit(10000)
        # canonical-hash guided descent: find a differing subtree cheaply
        d = None
        try:
            gview = m.words_np()
            # descend only where the subtree hashes differ
            def walk(gp, op_, level, path):
                leaf = level == NL - 1
                a, b = node(gview, gp, leaf), node(oview, op_, leaf)
                if leaf:
                    return None if a == b else (level, path, a, b)
                if a[0] != b[0]:
                    return (level, path, a, b)
                for i, (ca, cb) in enumerate(zip(a[1:], b[1:])):
                    ha = O.canonical(m.words_ptr, NL - level - 1, ca)["hash"] if False else None
                    r = walk(ca, cb, level + 1, path + [i])
                    if r:
                        return r
                return None
            d = walk(root, oroot, 0, [])
        except RecursionError:
            d = "recursion"
        print("MISMATCH it", it, "voxels", got["voxels"] - exp["voxels"], "by_ptr", got["by_ptr"], exp["by_ptr"], "levels",
              [a - b for a, b in zip(got["per_level"], exp["per_level"])], "filled", dev.FilledNodes(), "first diff", d, flush=True)
    dev.close()
print("done", N, "iterations,", bad, "mismatches")
