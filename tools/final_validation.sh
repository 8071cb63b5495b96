# final validation of a round: GPU tests, the C3 bench line, one traced C4 run
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_tests.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
MB2_RANSAC_TRACE=1 MB2_VERIFY_TRACE=1 python bench.py --workload c4 --steps 2 --warmup 3 > gpurun_out/final_c4_n1.json 2> gpurun_out/final_c4_n1.err
cat gpurun_out/final_tests.log; grep "verify\|mb2_ransac_h" gpurun_out/final_c4_n1.err | tail -2
python -c "
import json; d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1]); print('c3', d['value'], d['e2e']['value'], d['latency_ms_per_pair'], d['roofline']['frac'], d['roofline']['traffic'], d['parity']['ok'], d['stage_ms'])
d=json.loads(open('gpurun_out/final_c4_n1.json').read().strip().splitlines()[-1]); print('c4', d['value'], d['e2e']['value'], d['config']['single_pair'], d['config']['verify_ms_rank0'], d['config']['digest'][0])"
