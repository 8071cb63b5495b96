python tools/e2e_probe.py 10 2 > gpurun_out/s3_e2e_probe2.log 2>&1
python bench.py > gpurun_out/s3_bench2.json 2> gpurun_out/s3_bench2.err
cat gpurun_out/s3_e2e_probe2.log; python -c "
import json; d=json.loads(open('gpurun_out/s3_bench2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['clocks'], d['parity']['ok'])"
