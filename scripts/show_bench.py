#!/usr/bin/env python3
"""print the interesting numbers of a bench.py result line (last JSON line of the given file)"""
import json
import sys


def show(b, name):
    print("== %s value %.4g e2e %.4g (%.3f) ms %.3f frac %.3f kernel_ms %.3f results %s" % (
        name, b["value"], b["e2e"]["value"], b["e2e"]["value"] / b["value"], b["ms_per_step"],
        b["roofline"]["frac"], b["roofline"]["kernel_ms"],
        b.get("results_last_step", b.get("config", {}).get("results_last_step"))))
    print("   parallelism:", b.get("parallelism", b.get("config", {}).get("parallelism")))
    print("   parity", b["parity_check"])
    for k, l in b["legs"].items():
        print("   leg %s value %.4g frac %.3f results %d flagged %d" % (
            k, l["value"], l["roofline_frac"], l["results_last_step"], l["flagged_last_step"]),
            {a: round(v, 4) for a, v in l["phases_ms_per_step_rank0"].items()})
    for k, p in (b.get("patterns") or {}).items():
        print("   pat", k, json.dumps(p)[:330])


for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print("#", path, "n_gpus", d.get("n_gpus"))
    if d.get("impl") == "reference":
        print(json.dumps(d)[:500])
        continue
    show(d, "headline")
    for k, b in (d.get("secondary") or {}).items():
        if "unavailable" in b:
            print(k, b)
        else:
            show(b, k)
    for k in ("load", "cpu_baseline", "clocks"):
        if k in d:
            print(k, d[k])
