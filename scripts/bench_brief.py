import json, sys
d = json.load(open(sys.argv[1]))
print("value %.3e  ms/step %.1f  e2e %.3e (%.1f ms)  launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"] if d.get("e2e") else 0, d["e2e"]["ms_per_step"] if d.get("e2e") else 0, d["gpu_launches"]))
for k, v in d["kernels"].items():
    if v and "launches" not in v: print("  %-11s %.3f ms  %.0f GB/s" % (k, v["ms"], v["achieved_gbs"]))
    elif v and k == "pop_like" and "ms_per_call_with_aux" in v: print("  %-9s n=%4d  %8.3f ms/launch  %7.1f GB/s (%.3f of HBM)  %.3e units/s | with aux kernels %.3f ms/call (%.3f of HBM)" % (k, v["launches"], v["ms_per_launch"], v["achieved_gbs"], v["hbm_frac"], v["units_per_s"], v["ms_per_call_with_aux"], v["hbm_frac_with_aux"]))
    elif v: print("  %-9s n=%4d  %8.3f ms/launch  %7.1f GB/s (%.3f of HBM)  %.3e units/s" % (k, v["launches"], v["ms_per_launch"], v["achieved_gbs"], v["hbm_frac"], v["units_per_s"]))
print("  issue:", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["roofline"]["issue"].items()}, "share", round(d["roofline"]["share_of_step"], 3))
if d.get("cpu_baseline"): print("  cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["kind"])
print("  clocks:", d["clocks"])
