set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -n 5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 --no-dynamic > gpurun_out/bench_r2q_n8.json 2> gpurun_out/bench_r2q_n8.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_r2q_n8.json').read().strip().splitlines()[-1])
f=d['forward']
print('N8 weak ms', d['ms_per_step'], 'value', d['value'], 'strong', d['strong_scaling']['ms_per_step'], 'list', d['with_global_list']['ms_per_step'], 'verified', d['config']['exchange_verified_against_nccl'])
print('N8 forward', f['frame_ms'], f['shaded_mfrag_s'], f['split_equals_single_gpu_frame'], f['gpu_launches_per_frame'])
print({k:v.get('kernel_ms_per_frame') for k,v in f['roofline'].items() if isinstance(v,dict) and 'kernel_ms_per_frame' in v}, f['roofline']['other_stage_ms_per_frame'])
P
tail -n 8 gpurun_out/bench_r2q_n8.err
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 4 --steps 20 --warmup 5 --no-dynamic > gpurun_out/bench_r2q_n4.json 2> gpurun_out/bench_r2q_n4.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_r2q_n4.json').read().strip().splitlines()[-1])
f=d['forward']
print('N4 weak ms', d['ms_per_step'], 'value', d['value'], 'strong', d['strong_scaling']['ms_per_step'], 'list', d['with_global_list']['ms_per_step'], 'verified', d['config']['exchange_verified_against_nccl'])
print('N4 forward', f['frame_ms'], f['shaded_mfrag_s'], f['split_equals_single_gpu_frame'], f['gpu_launches_per_frame'])
P
