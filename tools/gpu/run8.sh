set -x
O=gpurun_out/r02h
mkdir -p $O
export NCCL_DEBUG=WARN
for cfg in c2 c1 c5; do
  T0=$SECONDS
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config $cfg --steps 4 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/bench2_$cfg.json 2> $O/bench2_$cfg.err
  echo "rc=$? wall=$((SECONDS-T0))s" >> $O/bench2_$cfg.err
  tail -c 400 $O/bench2_$cfg.json; tail -c 600 $O/bench2_$cfg.err
done
