import sys, numpy as np
sys.path.insert(0,'/root/repo')
from imageclust_b200 import clustering, _lib
rng = np.random.default_rng(11)
base = (rng.standard_normal((6000, 8)) * 50).astype(np.float32)
x = np.concatenate([base, base])[rng.permutation(12000)]
eng = clustering.Engine(0)
eng.set_option("gram_mode", 1)
for vr in (1,2):
  for near in (0,1):
    for comp in (0,1):
        eng.set_option("virtual_ranks", vr); eng.set_option("near_lists", near); eng.set_option("compact", comp)
        try:
            res = eng.cluster(x, 1, 4)
            st = res.stats
            print(vr, near, comp, "ok merges", st["n_merges"], "iters", st["n_iterations"], "raises", st["n_horizon_raises"], "comp", st["n_compactions"], flush=True)
        except Exception as e:
            print(vr, near, comp, "FAIL", str(e)[:150], flush=True)
            eng.close(); eng = clustering.Engine(0); eng.set_option("gram_mode", 1)
