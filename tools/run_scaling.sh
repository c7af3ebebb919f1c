#!/bin/bash
# One box, N GPUs: bench.py weak + strong, configs 2 and 5 (tools/bench_scaling.py). Usage: tools/run_scaling.sh N
N=$1
OUT=gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/r02_scale_weak_n1.json 2> $OUT/r02_scale_weak_n1.err
  cp $OUT/r02_scale_weak_n1.json $OUT/r02_scale_strong_n1.json
  python tools/bench_scaling.py $OUT/r02_scaling_cfg_n1.json > $OUT/r02_scaling_cfg_n1.log 2>&1
else
  L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  $L --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --no-parity-check > $OUT/r02_scale_weak_n$N.json 2> $OUT/r02_scale_weak_n$N.err
  $L --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 --no-parity-check --scaling strong > $OUT/r02_scale_strong_n$N.json 2> $OUT/r02_scale_strong_n$N.err
  $L --master-port 29543 tools/bench_scaling.py $OUT/r02_scaling_cfg_n$N.json > $OUT/r02_scaling_cfg_n$N.log 2>&1
fi
for f in $OUT/r02_scale_weak_n$N.json $OUT/r02_scale_strong_n$N.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["n_gpus"], d["scaling"], "ms %.3f e2e %.3f value %.4g e2e %.4g graph %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["cuda_graph"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
tail -3 $OUT/r02_scaling_cfg_n$N.log | cut -c1-200
