for hb in 4 2 1; do echo "== head_final blocks per SM $hb"; LFD_B200_HF_BLOCKS=$hb python bench.py --steps 50 --warmup 5 --no-cpu-baseline --profile-ops 2>gpurun_out/ops_$hb.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['impl_detail']['side_branch_ctas'])"; grep head_final gpurun_out/ops_$hb.err | head -2; done
python -m pytest tests/test_gpu_forward.py -q -m gpu -x 2>&1 | tail -2
