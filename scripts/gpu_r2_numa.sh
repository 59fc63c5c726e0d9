# NUMA-local pinned buffers (dist.bind_near_gpu): the weak-scaling line with its end-to-end block, bound and unbound
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for mode in bound unbound; do
  [ $mode = unbound ] && export WGS_NO_NUMA_BIND=1
  timeout 200 $TR --master-port 2963$N bench.py --gpus $N --steps 3 --no-extra --no-cpu-baseline > gpurun_out/r2_numa_${mode}_${N}gpu.json 2> gpurun_out/r2_numa_${mode}_${N}gpu.err; echo "$mode rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_numa_${mode}_${N}gpu.json"))
e=d["e2e"]
print("$mode", d["config"].get("host_binding_rank0"), "| step %.1f ms | e2e %.1f ms" % (d["ms_per_step"], e["ms_per_step"]), {k: e[k] for k in ("h2d_gbs_per_gpu", "upload_only") if k in e})
PY
done
