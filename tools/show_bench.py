import json, sys
l = [x for x in open(sys.argv[1]) if x.startswith("{")][-1]
d = json.loads(l)
print({k: round(d[k], 2) if isinstance(d[k], float) else d[k] for k in ("value", "ms_per_step", "gpu_launches")},
      "e2e", round(d["e2e"]["value"], 1), "full", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["full_channels"].items()})
r = d["roofline"]
print("roofline", r["kernel"], round(r["achieved"], 1), r["unit"], "frac", round(r["frac"], 3), "share", round(r["share_of_step"], 3))
print(d["kernel_breakdown_ms_per_step"]); print("dcs", d.get("kernel_breakdown_dcs_ms_per_step"))
print("attn TF", d.get("attn_tflops"), "clocks", d["clocks"], "cpu", d.get("cpu_baseline"))
print("eager", d.get("torch_eager_gpu"))
