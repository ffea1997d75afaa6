import json,sys
d=json.load(open(sys.argv[1]))
print("steps/s %.2f  ms/step %.2f  e2e %.2f  launches %d  inf ms/window %.2f e2e Mvox/s %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["inference"]["ms_per_window"], d["inference"]["e2e"]["value"]))
for k,v in d["roofline"]["families"].items(): print("  %-24s %6.2f ms  %3d launches  %7.1f GF  %s TF" % (k, v["ms_per_step"], v["launches_per_step"], v["gflop_per_step"], ("%.1f"%v["achieved_tflops"]) if v["achieved_tflops"] else "-"))
