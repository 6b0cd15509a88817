# N = 8 / 2 bench lines, with and without the NUMA binding of the ranks (run under gpurun --gpus 8)
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$T --nproc-per-node 8 --master-port 29533 bench.py --gpus 8 --steps 40 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
DATMO_NO_NUMA_BIND=1 $T --nproc-per-node 8 --master-port 29535 bench.py --gpus 8 --steps 40 --warmup 5 > gpurun_out/r02_bench_n8_nobind.json 2> gpurun_out/r02_bench_n8_nobind.err
$T --nproc-per-node 2 --master-port 29534 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)" >> gpurun_out/r02_topo.txt
for f in n8 n8_nobind n2; do python -c "import json,sys; d=json.loads(open('gpurun_out/r02_bench_$f.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['e2e']['value'], d.get('host_numa_binding_rank0'), d['shard_equality']['identical'])"; done
