#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
echo "rc=$?"; tail -5 gpurun_out/r2e_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/r2e_bench_n2_ref.json 2> gpurun_out/r2e_bench_n2_ref.err
echo "ref rc=$?"; cat gpurun_out/r2e_bench_n2_ref.json | cut -c1-400
LIFTREG_B200_ZERO_COPY=0 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2e_bench_staged.json 2> gpurun_out/r2e_bench_staged.err
python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2e_bench_zc.json 2> gpurun_out/r2e_bench_zc.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench_n2.json'))
for k in ('value','ms_per_step','n_gpus','e2e','cfg4_drr_view_sharded','cfg5_training_ops','sharded_parity','config'):
    print(k, json.dumps(d.get(k))[:1000])
for f in ('staged','zc'):
    d=json.load(open('gpurun_out/r2e_bench_%s.json'%f)); print(f, d['e2e']['ms_per_step'], d['e2e']['serial']['ms_per_step'], d['value'])
PY
timeout 600 python -m pytest tests/test_multigpu_nccl.py -m gpu -x -q 2>&1 | tail -3
