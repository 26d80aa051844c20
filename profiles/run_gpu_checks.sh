mkdir -p gpurun_out
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err); python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print(d['value'], d['e2e'], d['n_gpus'], d['batch'])"; tail -3 gpurun_out/bench_n2.err
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err); python -c "
import json; d=json.load(open('gpurun_out/bench_ref_n2.json')); print(d['value'], d['cpu_baseline'])"
