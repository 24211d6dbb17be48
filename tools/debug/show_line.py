"""Prints the headline fields of bench JSON lines.  usage: show_line.py file.json [...]"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "no line:", e); continue
    r = d.get("roofline") or {}
    print(f.split("/")[-1], "| value", round(d["value"]), d["unit"], "| ms/step", round(d["ms_per_step"], 4), "| e2e", round(d["e2e"]["value"]), "| n", d["n_gpus"], d["scaling"],
          "| parity", d.get("parity_check"), "| gather", d.get("gather_check"), "| launches", d.get("gpu_launches"))
    print("    pass_ms", {k: round(v, 3) for k, v in d["pass_ms"].items()}, "sum", round(d.get("pass_ms_sum", 0), 3), "| step_ms", {k: round(v, 3) for k, v in d["step_ms"].items() if k != "of"})
    print("    roofline", r.get("kernel"), "frac", round(r.get("frac", 0), 4), "launch ms", round(r.get("avg_launch_ms", 0), 4), "l2_gather_frac", round(r.get("l2_gather_frac", 0), 3),
          "| df_regen us", round(d["df_regen"]["us_per_regeneration"], 2) if d.get("df_regen") else None, "| cpu", (d.get("cpu_baseline") or {}).get("value"), "| clocks", d.get("clocks"))
