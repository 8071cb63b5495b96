python bench.py --workload c4 --steps 6 --warmup 3 > gpurun_out/s3_c4_ds_n1.json 2> gpurun_out/s3_c4_ds_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload c4 --steps 6 --warmup 3 > gpurun_out/s3_c4_ds_n2.json 2> gpurun_out/s3_c4_ds_n2.err; tail -3 gpurun_out/s3_c4_ds_n2.err
for f in s3_c4_ds_n1 s3_c4_ds_n2; do python -c "
import json; d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['single_pair'], d['config']['digest'], d['config']['digest_identical_on_all_ranks'], d['config']['verified'])"; done
