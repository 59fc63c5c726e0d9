# N-GPU call: the cfg4 target run, cfg3 weak + strong scaling lines (tag from $2)
N=${1:-8}; TAG=${2:-r2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29634 bench.py --gpus $N --config cfg4 --steps 3 > gpurun_out/${TAG}_cfg4_${N}gpu.json 2> gpurun_out/${TAG}_cfg4_${N}gpu.err; echo "cfg4 rc=$?"; python scripts/bench_brief.py gpurun_out/${TAG}_cfg4_${N}gpu.json 2>/dev/null | head -16; tail -2 gpurun_out/${TAG}_cfg4_${N}gpu.err
timeout 600 $TR --master-port 29635 bench.py --gpus $N --steps 5 --no-extra > gpurun_out/${TAG}_cfg3_weak_${N}gpu.json 2> gpurun_out/${TAG}_cfg3_weak_${N}gpu.err; echo "cfg3 weak rc=$?"; python scripts/bench_brief.py gpurun_out/${TAG}_cfg3_weak_${N}gpu.json 2>/dev/null | head -9
timeout 600 $TR --master-port 29636 bench.py --gpus $N --steps 5 --scaling strong --no-extra > gpurun_out/${TAG}_cfg3_strong_${N}gpu.json 2> gpurun_out/${TAG}_cfg3_strong_${N}gpu.err; echo "cfg3 strong rc=$?"; python scripts/bench_brief.py gpurun_out/${TAG}_cfg3_strong_${N}gpu.json 2>/dev/null | head -9
python -c "
import json
for f in ['cfg3_weak','cfg3_strong']:
    d=json.load(open('gpurun_out/${TAG}_%s_${N}gpu.json'%f)); print(f, 'e2e', {k:v for k,v in d['e2e'].items() if k in ('ms_per_step','h2d_gbs_per_gpu','upload_only')})
"
