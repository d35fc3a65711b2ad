nvidia-smi topo -m 2>&1 | head -20
lscpu | grep -i "numa\|socket\|^CPU(s)\|model name"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
RPO_BIND_NUMA=0 timeout 300 $TR --master-port 29631 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/b2_nobind.json 2>gpurun_out/b2_nobind.err
timeout 300 $TR --master-port 29632 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/b2_bind.json 2>gpurun_out/b2_bind.err
python - <<'PY'
import json
for f in ('gpurun_out/b2_nobind.json','gpurun_out/b2_bind.json'):
    t=open(f).read()
    print(f, 'lines', len(t.strip().splitlines()))
    d=json.loads(t.strip().splitlines()[-1]); print(d['value'], d['e2e'])
PY
