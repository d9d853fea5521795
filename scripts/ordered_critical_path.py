"""Length of the longest dependency chain of the ORDERED o2 update stream (no GPU needed).

A pair update reads and writes the node row of walk[j], the context row of walk[i] and `negative` sampled context rows
(pyx:105-151); it must see every earlier write of those rows.  level(pair) = 1 + max(level of the previous pair of the
walk, level of the latest earlier pair touching any of its rows); the largest level is the number of strictly
sequential steps ANY exact schedule needs, and pairs / longest = the parallelism an exact replay can use.
Walk tokens and samples are uniform here (the SBM corpus of BASELINE.json is close to that); hubs make it worse.

    python scripts/ordered_critical_path.py            # rows 1e5, 1e6, 1e7; 300 walks of 80, window 5, negative 5
"""
import sys

import numpy as np


def critical_path(rows, walks, length=80, window=5, neg=5, seed=1):
    rng = np.random.RandomState(seed)
    lvl_node = np.zeros(rows, np.int64)
    lvl_ctx = np.zeros(rows, np.int64)
    total = longest = 0
    for _ in range(walks):
        path = rng.randint(0, rows, size=length)
        prev = 0
        for i in range(length):
            wi = path[i]
            for j in range(max(0, i - window), min(length, i + window + 1)):
                if j == i:
                    continue
                wj = path[j]
                negs = rng.randint(0, rows, size=neg)
                lv = max(prev, lvl_node[wj], lvl_ctx[wi], lvl_ctx[negs].max()) + 1
                lvl_node[wj] = lv
                lvl_ctx[wi] = lv
                lvl_ctx[negs] = lv
                prev = lv
                total += 1
        longest = max(longest, prev)
    return total, longest


if __name__ == "__main__":
    for rows in [int(a) for a in sys.argv[1:]] or [100000, 1000000, 10000000]:
        t, l = critical_path(rows, 300)
        print("rows %9d  pairs %7d  longest chain %7d  parallelism %.2f" % (rows, t, l, t / l), flush=True)
