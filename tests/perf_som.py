"""(Not a pytest module: a timing script that uses the oracle as the CPU side, hence under tests/.)
SOM training at the reference's scale (refine.py:1043, 1066-1071: side = max(100, 5*bins),
800 iterations, <= 1000 vectors per iteration): device trainer vs the oracle on a few iterations."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import oracle
from groopm_b200 import engine, synth

side, n_train, iters = 100, 5000, 800
nodes = synth.som_nodes(side, 4, 1)
block, _ = synth.cfg_bmu(n_train, 8, 2)
np.random.seed(1)
engine.som_train(nodes, block, iterations=5, influence_rate=0.15)          # warm-up (library load, allocator)
np.random.seed(1)
t0 = time.perf_counter()
w = engine.som_train(nodes, block, iterations=iters, influence_rate=0.15)
t_dev = time.perf_counter() - t0
done = int(np.floor(np.log(50.0 / 2.0) * iters / np.log(50.0)))          # iterations before the radius drops under 2
print(f"device: side {side}, {n_train} vectors, {iters} iterations requested ({done} run): {t_dev:.2f} s "
      f"= {t_dev / done * 1e3:.2f} ms / iteration")
np.random.seed(1)
t0 = time.perf_counter()
w_o = oracle.som_train(nodes, block, iterations=iters, influence_rate=0.15, vector_subset=1000) if "--full" in sys.argv else None
if w_o is None:
    # the oracle's cost per iteration is flat in the iteration number: time 3 iterations of the same schedule
    import types
    t0 = time.perf_counter()
    np.random.seed(1)
    # run the first 3 iterations only by asking for the same decay schedule and stopping early
    class Stop(Exception):
        pass
    cnt = [0]
    orig = np.random.shuffle
    def counting_shuffle(a):
        cnt[0] += 1
        if cnt[0] > 3:
            raise Stop()
        return orig(a)
    np.random.shuffle = counting_shuffle
    try:
        oracle.som_train(nodes, block, iterations=iters, influence_rate=0.15)
    except Stop:
        pass
    finally:
        np.random.shuffle = orig
    t_cpu3 = time.perf_counter() - t0
    print(f"oracle (scipy/numpy, 1 thread): {t_cpu3 / 3:.2f} s / iteration -> {t_cpu3 / 3 * done:.0f} s for {done} iterations (extrapolated); "
          f"speed-up {t_cpu3 / 3 * done / t_dev:.0f}x")
else:
    print("oracle full:", time.perf_counter() - t0, "s; identical:", np.array_equal(w, w_o))
