# bench at 2 GPUs: the closing mode of a sharded batch (1 = own exchange kernel, 0 = last item's kernel)
for mode in ${MODES:-1 0}; do
  for rep in 1 2; do
    CSGN_TUNING=1 CSGN_PEER_CLOSE_KERNEL=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 2951$rep bench.py --gpus ${NG:-2} --steps 100 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('close_kernel=$mode', 'value %.4g ms %.4f e2e %.4g ms %.4f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))"
  done
done
