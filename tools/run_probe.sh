python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/s3_c4_n1.json 2> gpurun_out/s3_c4_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload c4 --steps 3 --warmup 3 > gpurun_out/s3_c4_n2.json 2> gpurun_out/s3_c4_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/s3_c3_n2.json 2> gpurun_out/s3_c3_n2.err
for f in s3_c4_n1 s3_c4_n2 s3_c3_n2; do python -c "
import json,sys; d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('digest'), d['config'].get('rank0_ms'))"; done
