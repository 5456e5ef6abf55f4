"""Print the headline numbers and the per-op table of a bench.py JSON line (development aid)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.0f %s  ms/step %.4f  e2e %.0f  graph kernels %s" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"],
                                                                  d["config"].get("graph_kernels_per_step")))
print("roofline", d["roofline"])
for k in d:
    if k.startswith("op_table"):
        for r in d[k]:
            print("  ", r)
