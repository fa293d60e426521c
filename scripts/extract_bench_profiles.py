"""Copies the bench lines of a round's gpurun runs into profiles/ (tracked): the whole line per N, plus the C4 / C5 blocks on
their own (profiles/r2_c4_sharded_n{N}.json, profiles/r2_c5_streams_n{N}.json) as the round-1 verdict asked.
  python scripts/extract_bench_profiles.py r2 gpurun_out/r2_bench_n1.json gpurun_out/r2_bench_n2.json ..."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
for path in sys.argv[2:]:
    line = json.loads(open(path).read().strip().splitlines()[-1])
    n = line["n_gpus"]
    json.dump(line, open(os.path.join(ROOT, "profiles", "%s_bench_n%d.json" % (tag, n)), "w"), indent=1)
    for key, name in (("C4", "c4_sharded"), ("C5", "c5_streams"), ("C3", "c3_device")):
        blk = (line.get("configs") or {}).get(key)
        if blk:
            blk = dict(blk, n_gpus=n, source="bench.py configs block, %s" % os.path.basename(path))
            json.dump(blk, open(os.path.join(ROOT, "profiles", "%s_%s_n%d.json" % (tag, name, n)), "w"), indent=1)
    print(n, round(line["value"]), round(line["e2e"]["value"]))
