#!/bin/bash
# One column of the result table: every config on N GPUs of this box.
#   tools/scale_table.sh N [steps]        -> gpurun_out/table/<config>_n<N>[_variant].json
N=${1:-1}
STEPS=${2:-10}
OUT=gpurun_out/table
mkdir -p $OUT
run() {  # name, args...
    local name=$1; shift
    if [ "$N" -gt 1 ]; then
        python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
            bench.py --gpus $N --steps $STEPS --no-cpu "$@" > $OUT/${name}_n$N.json 2> $OUT/${name}_n$N.err
    else
        python bench.py --gpus 1 --steps $STEPS "$@" > $OUT/${name}_n$N.json 2> $OUT/${name}_n$N.err
    fi
    echo "$name n=$N rc=$? $(python - <<PY
import json
try:
    d = json.loads(open("$OUT/${name}_n$N.json").read().strip().splitlines()[-1])
    g = d.get("gather") or {}
    print("value %.3e  ms %.3f  e2e ms %.3f  frac %s  gather_exposed %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"],
          (d.get("roofline") or {}).get("frac"), g.get("exposed_ms")))
except Exception as e:
    print("unreadable:", e)
PY
)"
}
run c2
if [ "$N" -gt 1 ]; then
    run c2_gather_p2p --gather p2p
    run c2_gather_nccl --gather nccl
fi
run c3 --workload c3 --steps 3
run c4 --workload c4
run c5 --workload c5
run c5cond --workload c5cond --steps 32
run krige --workload krige --steps 3
if [ "$N" -gt 1 ]; then
    python bench.py --plan $N --steps $STEPS --no-cpu > $OUT/c2_plan_n$N.json 2> $OUT/c2_plan_n$N.err
    echo "c2 single-process plan n=$N rc=$? $(python -c "
import json
d = json.loads(open('$OUT/c2_plan_n$N.json').read().strip().splitlines()[-1])
print('value %.3e  ms %.3f  e2e ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step']))")"
    python tools/d2h_probe.py > $OUT/d2h_probe_n$N.log 2>&1
    grep own_pinned $OUT/d2h_probe_n$N.log | tail -8
fi
