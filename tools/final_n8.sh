#!/bin/bash
# Final 8-GPU confirmation of HEAD: multi-GPU tests, host D2H probe (pinned / write-combined / NUMA-interleaved),
# config 2 / 3 under torchrun (plain step + p2p gather measured beside it), the C5 ensemble from one process.
OUT=gpurun_out/final_n8
mkdir -p $OUT
python -m pytest tests/test_gpu_multi.py tests/test_gpu_plan.py -x -q > $OUT/pytest_multi.log 2>&1; echo "multi tests rc=$?"; tail -2 $OUT/pytest_multi.log
python tools/d2h_probe.py > $OUT/d2h_probe.log 2>&1; grep -v h2d $OUT/d2h_probe.log | grep '"gpus": 8\|numa'
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
$TR bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/c2_n8.json 2> $OUT/c2_n8.err; echo "c2 rc=$?"
$TR bench.py --gpus 8 --steps 3 --workload c3 --no-cpu > $OUT/c3_n8.json 2> $OUT/c3_n8.err; echo "c3 rc=$?"
$TR bench.py --gpus 8 --steps 32 --workload c5cond > $OUT/c5cond_n8.json 2> $OUT/c5cond_n8.err; echo "c5cond rc=$?"
python bench.py --workload c5ens --steps 3 --plan 8 > $OUT/c5ens_plan8.json 2> $OUT/c5ens_plan8.err; echo "c5ens rc=$?"
python bench.py --plan 8 --steps 10 --no-cpu > $OUT/c2_plan8.json 2> $OUT/c2_plan8.err; echo "plan rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final_n8/*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        g = d.get("gather") or {}
        print(f.split("/")[-1], "value %.3e ms %.3f e2e ms %.3f frac %s gather %s" % (d["value"], d["ms_per_step"],
              d["e2e"]["ms_per_step"], (d.get("roofline") or {}).get("frac"), g.get("exposed_ms")))
    except Exception as e:
        print(f, "unreadable", e)
PY
