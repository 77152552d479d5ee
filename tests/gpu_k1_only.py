import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import load_pkg
k = load_pkg()
big = k.Problem.synth(4, 100000, 1000, 42, 0)
rb = k.ResidentSolve(big); rb.load()
print(sorted(rb.run_feasibility(flush_l2=True) for _ in range(6)))
