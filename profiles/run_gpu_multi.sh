mkdir -p gpurun_out
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err); wc -l gpurun_out/bench_n4.json; python -c "
import json; d=json.load(open('gpurun_out/bench_n4.json')); print(d['value'], d['e2e']['value'], d['n_gpus'], d['batch']['qp_iterations_per_sec'], d['batch']['qps_per_gpu'])"
