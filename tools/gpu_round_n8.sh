#!/bin/bash
# 8-GPU call: (1) bench.py under torchrun at N=8 (headline cfg2 weak scaling + extras: cfg4 / cfg5 / cfg3 strong, cfg3 with oracle parity),
# (2) one host process driving 1/2/4/8 GPUs through sage_b200_score_batch_multi, (3) the multi-GPU parity test, (4) topology.
cd "$(dirname "$0")/.."
tag=${1:-r02_n8}; out=gpurun_out/$tag; mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
nproc >> $out/topo.txt; numactl -H >> $out/topo.txt 2>&1
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 3 ) > $out/bench_n8.json 2> $out/bench_n8.err
tail -5 $out/bench_n8.err
python - $out/bench_n8.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1]); p=d["phases_ms_per_step"]
    print("N=8 value %.2fM e2e %.2fM (%.2f ms) pageable %.2fM | prelim %.3f score %.3f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["e2e"].get("pageable",{}).get("value",0)/1e6, p["prelim"], p["score"]), d.get("numa_node"))
    for k,v in d.get("extra",{}).items(): print("   extra", k, ("value %.3fM e2e %.3fM" % (v["value"]/1e6, v["e2e"]["value"]/1e6)) if "value" in v else v, v.get("parity_checked"), v.get("index"))
except Exception as e: print("FAILED", e)
PY
( time timeout 900 python tools/bench_multi.py --steps 10 ) > $out/bench_multi.json 2> $out/bench_multi.err
tail -2 $out/bench_multi.json | cut -c1-1500
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi" > $out/test_multi.log 2>&1; tail -2 $out/test_multi.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --steps 20 --warmup 3 --no-extras ) > $out/bench_n4.json 2> $out/bench_n4.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 3 --no-extras ) > $out/bench_n2.json 2> $out/bench_n2.err
for n in 2 4; do python - $out/bench_n$n.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    print(sys.argv[1], "value %.2fM e2e %.2fM" % (d["value"]/1e6, d["e2e"]["value"]/1e6))
except Exception as e: print("FAILED", e)
PY
done
