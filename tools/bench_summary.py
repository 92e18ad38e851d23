import json, sys
l = [x for x in open(sys.argv[1]) if x.startswith("{")]
if not l:
    print(open(sys.argv[1]).read()[-2000:]); sys.exit(0)
d = json.loads(l[-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]), d["clocks"], "launches", d["gpu_launches"])
print([(s["name"], s["ms"], s["tflops"], s["gbs"]) for s in d["roofline"]["stages"]])
r = d["roofline"]
print("dominant", r["kernel"], r["bound"], round(r["achieved"], 1), r["unit"], "frac", round(r["frac"], 3), "share", round(r["share_of_step"], 3))
print([(k["name"], k["n"], k["ms"], k["tflops"], k["gbs"]) for k in r.get("kernels", [])])
if "cpu_baseline" in d: print(d["cpu_baseline"])
