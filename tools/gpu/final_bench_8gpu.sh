P=gpurun_out/r2fin4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 30 --warmup 5 --no-extras > ${P}_bench_c2_n8.json 2> ${P}_bench_c2_n8.err; echo "bench n8 rc=$?"
python -c "
import json
d=json.loads([l for l in open('${P}_bench_c2_n8.json') if l.startswith('{')][-1])
print(d['value']/1e9, d['ms_per_step'], d['parity']['equal']); print(d['dist_stage_ms']); print(d['e2e']['ms_per_step'])"
