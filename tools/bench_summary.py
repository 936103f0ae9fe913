import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f)); r=d["roofline"]
        print(f, "value",round(d["value"]), "ms/step", round(d["ms_per_step"],4), "fe_ms", round(r["avg_launch_ms"],4), "fe_frac", round(r["frac"],3), "rec_ms", round(r["recurrent_kernel"]["avg_launch_ms"],4), "whole", round(r["whole_step"]["frac"],3), "e2e", round(d["e2e"]["value"]), "p99", round(d["p99_step_ms"],4), d["clocks"])
    except Exception as e: print(f, "ERR", e)
