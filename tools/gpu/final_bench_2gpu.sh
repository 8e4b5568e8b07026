P=gpurun_out/r2fin2
python -m pytest tests/test_fullsize.py -q -m gpu -k nccl 2>&1 | tail -3 | tee ${P}_pytest_nccl2.log
for i in 1 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras > ${P}_bench_c2_n2.json 2> ${P}_bench_c2_n2.err; echo "bench n2 rc=$?"
python -c "
import json
d=json.loads([l for l in open('${P}_bench_c2_n2.json') if l.startswith('{')][-1])
print(d['value']/1e9, d['ms_per_step'], d['parity']['equal']); print(d['dist_stage_ms'])"
done
