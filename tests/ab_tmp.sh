python -m pytest tests/test_gpu_train.py tests/test_gpu_executor.py tests/test_gpu_ops.py tests/test_gpu_variants.py -q -m gpu 2>&1 | tail -4
python bench.py --config WIDERFACE_L_train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/train_lazy.json 2> gpurun_out/train_lazy.err; tail -2 gpurun_out/train_lazy.err
python -c "
import json
d=json.loads(open('gpurun_out/train_lazy.json').read().strip().splitlines()[-1]); print('train', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['impl_detail']['block_ms'], d['impl_detail']['loss_first_last'])"
