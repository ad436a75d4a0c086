#!/bin/bash
set +e
P='import sys,json; d=json.loads(sys.stdin.read()); print({k: d.get(k) for k in ("value","ms_per_step","n_gpus","host_numa")}, "e2e", d["e2e"] and d["e2e"]["value"], "train", d["train_step"] and d["train_step"]["ms_per_step"], "clocks", d["clocks"])'
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 8000 --warmup 100 2> gpurun_out/n2.err | tail -1 | python -c "$P" || tail -20 gpurun_out/n2.err
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>> gpurun_out/n2.err | tail -1 | cut -c1-600
