N=$1
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q -rs > gpurun_out/r2_sharded_${N}gpu.log 2>&1; tail -5 gpurun_out/r2_sharded_${N}gpu.log
for C in ${CONFIGS:-4 3 5 2}; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $C --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_config${C}_${N}gpu.json 2> gpurun_out/r2_config${C}_${N}gpu.err; tail -c 600 gpurun_out/r2_config${C}_${N}gpu.json; tail -2 gpurun_out/r2_config${C}_${N}gpu.err
done
