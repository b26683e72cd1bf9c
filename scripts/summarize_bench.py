"""Developer helper: the headline keys of bench.py JSON lines (files given on the command line)."""
import json
import sys
for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    mg = d.get("multi_gpu") or {}
    print(f, "| N", d["n_gpus"], "| value", d["value"], "| ms/step", d["ms_per_step"], "| e2e", d["e2e"]["value"], "| pageable", d["e2e"]["pageable"]["value"],
          "| link", d["e2e"].get("link", {}).get("frac_of_link"), "| path trace", {k: v["msamples_s"] for k, v in d["path_trace"].items()},
          "| checks", {k: v for k, v in (mg.get("checks") or {}).items() if isinstance(v, bool)}, "| strong", (mg.get("strong_scaling") or {}).get("mrays_s"))
