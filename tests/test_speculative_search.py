"""The sharded consolidation search must replay exactly the sequential binary search of
multinodeconsolidation.go:74-114 whatever the number of ranks. CPU only: the probe is a table lookup here; the GPU
tests run the same search with real simulateScheduling probes against the oracle."""
import os
import random
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def sequential(n, table):
    """the reference loop, restated: min=1, max=n-1, probe candidates[0:mid+1]"""
    if n < 2:
        return (0, 0, []), []
    lo, hi, last, path = 1, n - 1, (0, 0, []), []
    while lo <= hi:
        mid = (lo + hi) // 2
        action, options = table[mid + 1]
        path.append(mid + 1)
        if action in (1, 2):
            last = (action, mid + 1, options)
            lo = mid + 1
        else:
            hi = mid - 1
    return last, path


def random_table(rng, n, monotone):
    if monotone:
        cut = rng.randint(0, n)
        acts = [rng.choice((1, 2)) if c <= cut else 0 for c in range(n + 1)]
    else:
        acts = [rng.choice((0, 1, 2)) for _ in range(n + 1)]
    return {c: (acts[c], [c, c + 1] if acts[c] == 2 else []) for c in range(n + 1)}


@pytest.mark.parametrize("width", [1, 2, 3, 4, 7, 8, 15, 64])
def test_speculative_search_equals_sequential(pkg, width):
    rng = random.Random(width)
    for trial in range(300):
        n = rng.choice([0, 1, 2, 3, 4, 5, 9, 17, 33, 100, 101])
        table = random_table(rng, n, monotone=trial % 2 == 0)
        calls = []

        def probe_many(counts):
            assert len(counts) <= width and len(set(counts)) == len(counts)
            assert all(2 <= c <= n for c in counts)
            calls.append(list(counts))
            return {c: table[c] for c in counts}

        action, count, options, rounds, path = pkg.speculative_binary_search(n, probe_many, width)
        want, want_path = sequential(n, table)
        assert (action, count, options) == want
        assert path == want_path
        assert rounds == len(calls)
        if n >= 2 and width >= 1:
            import math
            depth = math.floor(math.log2(width + 1))
            assert rounds <= math.ceil(len(want_path) / depth) + 1


def test_frontier_is_breadth_first_and_whole_levels(pkg):
    assert pkg.speculation_frontier(1, 99, 1) == [51]
    f3 = pkg.speculation_frontier(1, 99, 3)
    assert f3[0] == 51 and set(f3[1:]) == {76, 26}
    f7 = pkg.speculation_frontier(1, 99, 7)
    assert len(f7) == 7 and len(set(f7)) == 7
    assert pkg.speculation_frontier(5, 4, 8) == []
    assert pkg.speculation_frontier(3, 3, 8) == [4]


WORKER = textwrap.dedent("""
    import os, sys, random
    sys.path.insert(0, {root!r})
    import torch.distributed as dist
    import __graft_entry__ as g
    pkg = g.load_pkg()
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = random.Random(7)
    n = 100
    acts = [rng.choice((1, 2)) if c <= 61 else 0 for c in range(n + 1)]
    evaluated = []
    def probe_many(counts):
        mine = {{}}
        for i, c in enumerate(counts):
            if i % world == rank:
                evaluated.append(c)
                mine[c] = (acts[c], [c] if acts[c] == 2 else [])
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        merged = {{}}
        for part in parts:
            merged.update(part)
        return merged
    out = pkg.speculative_binary_search(n, probe_many, world)
    print("RESULT", rank, out[0], out[1], out[3], len(evaluated), flush=True)
    dist.destroy_process_group()
""")


def test_world_size_two_gloo_probes_are_split_and_answer_is_sequential(tmp_path):
    port = 29600 + os.getpid() % 300
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT), port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    res = [[x for x in o.splitlines() if x.startswith("RESULT")][0].split() for o in outs]
    # both ranks agree, answer = the sequential one (largest consolidatable prefix: 61 candidates)
    assert res[0][2:5] == res[1][2:5]
    assert int(res[0][3]) == 61 and int(res[0][2]) in (1, 2)
    # the two ranks evaluated disjoint halves of each round's frontier
    assert int(res[0][5]) + int(res[1][5]) <= 2 * int(res[0][4])
