import json, sys
d = json.load(open(sys.argv[1]))
e = d.get("e2e") or {}
print("value %.3e  ms/step %.1f  e2e %.3e (%.1f ms)  launches %d  n_gpus %d  %s" % (d["value"], d["ms_per_step"], e.get("value", 0), e.get("ms_per_step", 0), d["gpu_launches"], d["n_gpus"], d["config"]["workload"][:40]))
for k, v in d["kernels"].items():
    if not v: continue
    if k == "ingest":
        for kk, vv in v.items():
            if isinstance(vv, dict): print("  ingest %-16s %7.1f MB/s compressed  %7.1f MB/s text  wall %.2f s" % (kk, vv.get("compressed_mb_per_s", 0), vv.get("uncompressed_mb_per_s", 0), vv.get("wall_s", 0)))
            else: print("  ingest", kk, vv)
        continue
    if "launches" not in v: print("  %-14s %.3f ms  %.0f GB/s" % (k, v["ms"], v["achieved_gbs"]))
    else:
        extra = " | with aux %.3f ms/call (%.3f of HBM)" % (v["ms_per_call_with_aux"], v["hbm_frac_with_aux"]) if "ms_per_call_with_aux" in v else ""
        print("  %-14s n=%5d  %8.3f ms/launch  %9.3f ms total  %7.1f GB/s (%.3f of HBM)  %.3e units/s%s" % (k, v["launches"], v["ms_per_launch"], v["ms_total"], v["achieved_gbs"], v["hbm_frac"], v["units_per_s"], extra))
r = d.get("roofline")
if r: print("  roofline: %s bound=%s frac=%.4f achieved=%.4g peak=%.4g share=%.3f" % (r["kernel"], r["bound"], r["frac"], r["achieved"], r["peak"], r["share_of_step"]))
if d.get("cpu_baseline"): print("  cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["kind"])
if d.get("parity"): print("  parity:", {k: v for k, v in d["parity"].items() if k != "against"})
print("  clocks:", d["clocks"])
if d.get("stream_idle"): print("  stream idle:", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["stream_idle"].items() if k != "note"})
