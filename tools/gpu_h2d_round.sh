#!/bin/bash
# concurrent H2D roofline at 1/2/4/8 ranks + e2e bench at the same rank counts (one 8-GPU box): profiles/r02_h2d_concurrent.txt
out=gpurun_out/r2_h2d_concurrent.txt
: > $out
for n in 1 2 4 8; do
  echo "== $n ranks: concurrent plain copies" >> $out
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) tools/gpu_h2d_concurrent.py 2>/dev/null | grep '^{' >> $out
done
for n in 1 2 4 8; do
  echo "== $n ranks: bench.py e2e" >> $out
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'n':d['n_gpus'],'value':d['value'],'e2e':d['e2e']}))" >> $out
done
nvidia-smi topo -m >> $out 2>&1
cat $out
