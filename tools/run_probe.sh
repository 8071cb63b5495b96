python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/s3_final_tests.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s3_final_ref.json 2> gpurun_out/s3_final_ref.err
python bench.py > gpurun_out/s3_final_bench.json 2> gpurun_out/s3_final_bench.err
bash tools/ncu_capture_r2.sh r2c > gpurun_out/s3_final_ncu.log 2>&1
cat gpurun_out/s3_final_tests.log; tail -c 600 gpurun_out/s3_final_ref.json; python -c "
import json; d=json.loads(open('gpurun_out/s3_final_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['latency_ms_per_pair'], d['roofline']['frac'], d['roofline']['ms_per_launch_group'], d['parity']['ok'], d['clocks'])"
tail -3 gpurun_out/s3_final_ncu.log
