import sys, json, fileinput
for ln in fileinput.input():      # files given as arguments, else stdin
    ln = ln.strip()
    if not ln.startswith("{"): continue
    d = json.loads(ln)
    r = d.get("roofline", {})
    print("%s: %.0f img/s  %.4f ms/step  frac %.3f  e2e %.0f  clocks %s  launches %s" % (
        d["config"]["workload"][:24], d["value"], d["ms_per_step"], r.get("frac", 0), d["e2e"]["value"],
        (d.get("clocks") or {}).get("sm_mhz"), d.get("gpu_launches")))
