# multi-GPU bench lines (run under gpurun --gpus N): N = number of visible GPUs
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err); wc -l gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_n$N.json')); print(json.dumps({k:d[k] for k in ('value','ms_per_step','scaling','n_gpus','e2e')})); print(json.dumps(d['batch'])); print(json.dumps(d.get('batch_weak'))); print(json.dumps(d.get('replicas_c2')))"
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err); cat gpurun_out/bench_ref_n$N.json | cut -c1-900
