mkdir -p gpurun_out/r02
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02/x_n1.json 2> gpurun_out/r02/x_n1.err
for n in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r02/x_n$n.json 2> gpurun_out/r02/x_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --workload c3 --sharding image --no-configs --steps 10 --warmup 3 > gpurun_out/r02/x_n8_c3_image.json 2> gpurun_out/r02/x_n8_c3_image.err
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.loads(open(f'gpurun_out/r02/x_n{n}.json').read().strip().split('\n')[-1])
    except Exception as e:
        print(n,'ERR',e); continue
    c3=d['configs'].get('c3',{})
    print(n, 'C5 ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3), 'C3 ms', round(c3.get('ms_per_step',0),3), 'e2e', round(c3.get('e2e',{}).get('ms_per_step',0),3), 'parity', d.get('parity_rel_l2_vs_reference'), c3.get('parity_rel_l2_vs_reference'), d['clocks'])
d=json.loads(open('gpurun_out/r02/x_n8_c3_image.json').read().strip().split('\n')[-1])
print('c3 image n8', d['ms_per_step'], d['e2e']['ms_per_step'])
PY
