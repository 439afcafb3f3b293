python profiles/microbench/rw_peaks.py 2>&1 | tail -4
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/b_dir.json 2> gpurun_out/b_dir.err; tail -2 gpurun_out/b_dir.err; python -c "
import json;d=json.loads(open('gpurun_out/b_dir.json').read().strip().splitlines()[-1]);n=d['roofline']['net'];print(round(d['value']),d['ms_per_step'],round(d['e2e']['value']),n['directional_peaks'],n['layerwise_bound_ms'],n['directional_bound_ms'],n['frac_of_layerwise_bound_in_graph'],n['frac_of_directional_bound_in_graph'])"
