#!/usr/bin/env python
"""Assert that every line of a .jsonl written by scripts/exp_cfg3_multigpu.py (or bler_sweep.py) for different world sizes
carries the same counters: the workload is geometry independent, so N GPUs must count exactly what one GPU counts."""
import json
import sys

lines = [json.loads(l) for l in open(sys.argv[1]) if l.strip().startswith("{")]
key = lambda d: json.dumps(d.get("counters") or [[p[k] for k in ("snr_db", "txBlocks", "tbCrcFail", "cbCrcFail", "bitErrors", "sumIterations")] for p in d["points"]], sort_keys=True)
groups = {}
for d in lines:
    cfg = d["config"] if isinstance(d["config"], str) else json.dumps({k: v for k, v in d["config"].items() if k != "out"}, sort_keys=True)
    groups.setdefault(cfg, []).append(d)
ok = True
for cfg, ds in groups.items():
    ks = {key(d) for d in ds}
    worlds = sorted({d["world"] for d in ds})
    print("%s\n   worlds %s -> %s" % (cfg[:150], worlds, "IDENTICAL counters" if len(ks) == 1 else "MISMATCH"))
    ok = ok and len(ks) == 1
sys.exit(0 if ok else 1)
