"""cProfile of the unchanged per-layer loop (tools/layer_api_bench.py): where the host time of a step goes."""
import cProfile, pstats, io, os, sys, runpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = ["layer_api_bench.py"]
import tools.layer_api_bench as lb   # noqa: runs the warm-up and the timed loop once
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    lb.step()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue()[:5000])
