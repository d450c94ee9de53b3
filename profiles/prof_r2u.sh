#!/bin/bash
# round 2, pass u: N GPUs (NG from the environment): sharding tests, the default bench line under torchrun (with the
# one-neighbourhood strong-scaling leg and BASELINE configs[3] -- random 128 x 1M trees sharded across the ranks)
NG=${NG:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharding_gpu.py -m gpu -q > gpurun_out/r2u_tests_n$NG.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2u_tests_n$NG.log
tail -4 gpurun_out/r2u_tests_n$NG.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 5 --warmup 3 > gpurun_out/bench_r2u_n$NG.json 2> gpurun_out/bench_r2u_n$NG.err ) 2>&1 | tail -3
tail -c 300 gpurun_out/bench_r2u_n$NG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2u_n$NG.json'))
print(d['n_gpus'], d['ms_per_step'], d['value'], d['roofline']['frac'])
print(d.get('one_neighbourhood_sharded'))
t=d.get('trees_c4'); print(t and (t['ms_per_step'], t['value'], t['trees_per_step']))
PY
