python -m pytest tests/test_gpu_train.py -q -m gpu -k ddp 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tests/run_train_ddp.py > gpurun_out/ddp_2gpu_final.log 2>&1; tail -4 gpurun_out/ddp_2gpu_final.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --config WIDERFACE_L_train --steps 10 --warmup 3 > gpurun_out/final2_train_2gpu.json 2> gpurun_out/final2_train_2gpu.err
python -c "
import json;d=json.loads(open('gpurun_out/final2_train_2gpu.json').read().strip().splitlines()[-1]);print('train 2gpu',round(d['value']),d['ms_per_step'],round(d['e2e']['value']),d['impl_detail'].get('allreduce_us'))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/final2_infer_2gpu.json 2> gpurun_out/final2_infer_2gpu.err
python -c "
import json;d=json.loads(open('gpurun_out/final2_infer_2gpu.json').read().strip().splitlines()[-1]);print('infer 2gpu',round(d['value']),d['ms_per_step'],round(d['e2e']['value']),d['e2e'].get('h2d_gbps_per_rank'))"
