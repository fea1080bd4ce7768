"""Where does the HOST time of one SP train step go?  cProfile over a few steps (the GPU runs ahead asynchronously)."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "egocentric-gaze-prediction_b200")); sys.path.insert(0, ROOT)
import torch
import bench
wl = bench.Workload(os.environ.get("WL", "sp_train"), 32, 224, 0, 1, torch.device("cuda"))
for _ in range(4):
    wl.step(*wl.dev)
torch.cuda.synchronize()
N = 5
t0 = time.perf_counter()
for _ in range(N):
    wl.step(*wl.dev)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host enqueue %.2f ms/step (unprofiled)" % ((t1 - t0) * 1e3 / N))
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    wl.step(*wl.dev)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
