python -m pytest tests -m gpu -x -q -k "ransac or fullsize or mods_pair or verif" 2>&1 | tail -2
MB2_RANSAC_TRACE=1 MB2_VERIFY_TRACE=1 python bench.py --workload c4 --steps 2 --warmup 1 2>gpurun_out/s3_c4_trace.err >gpurun_out/s3_c4_b.json; grep "verify\|mb2_ransac_h" gpurun_out/s3_c4_trace.err | tail -2
python -c "
import json; d=json.loads(open('gpurun_out/s3_c4_b.json').read().strip().splitlines()[-1]); print('c4', d['value'], d['ms_per_step'], d['config']['digest'], d['config']['rank0_ms'])"
python bench.py > gpurun_out/s3_bench4.json 2>gpurun_out/s3_bench4.err
python -c "
import json; d=json.loads(open('gpurun_out/s3_bench4.json').read().strip().splitlines()[-1]); print('c3', d['value'], d['e2e']['value'], d['latency_ms_per_pair'], d['stage_ms'], d['parity']['ok'])"
