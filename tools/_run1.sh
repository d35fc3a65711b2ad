TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29621 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu_h.json 2> gpurun_out/bench_8gpu_h.err; tail -c 600 gpurun_out/bench_8gpu_h.json | head -c 100; echo
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_8gpu_h.json'))
print({k:d[k] for k in ('value','n_gpus','ms_per_step','clocks')})
PY
timeout 300 $TR --master-port 29622 examples/train_evopf_dp.py --algo sac --graph --iters 300 2>&1 | tail -1 | tee gpurun_out/train_sac_8gpu.json
timeout 300 $TR --master-port 29623 examples/train_evopf_dp.py --graph --iters 300 2>&1 | tail -1 | tee gpurun_out/train_ddpg_8gpu.json
