python -m pytest tests/test_gpu_conv.py tests/test_gpu_forward.py tests/test_gpu_fullsize.py -q -m gpu -x 2>&1 | tail -4
python tests/debug_forward_timing.py WIDERFACE_S 8 720 1280 stem2deep 2>&1 | tail -1
python tests/debug_forward_timing.py WIDERFACE_XS 1 480 640 stem2deep 2>&1 | tail -1
python tests/debug_forward_timing.py WIDERFACE_XS 2 2160 3840 stem2deep 2>&1 | tail -1
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --profile-ops 2>gpurun_out/ops.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['impl_detail']['side_branch_ctas'], d['roofline']['kernel'], d['roofline']['frac'])"
head -4 gpurun_out/ops.err
