import sys, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import depecher_b200 as dp, depeche_c3
eng = dp.Engine(device=0)
for i in range(4):
    r = depeche_c3.run(eng, 10_000_000, 40, 30, 10, 24, 2.5, 20261021, verbose=False)
    print({k: r[k] for k in ("penalty_opt_s", "selection_s", "final_allocate_s", "total_s", "kernel_launches")}, flush=True)
