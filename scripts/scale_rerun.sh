run() { n=$1; tag=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+RANDOM%200)) bench.py --gpus $n "$@" > gpurun_out/r02_${tag}.json 2> gpurun_out/r02_${tag}.err
  echo "$tag rc=$? $(cut -c1-200 gpurun_out/r02_${tag}.json)"; }
run 2 scale_2gpu --steps 20 --warmup 5
run 4 scale_4gpu --steps 20 --warmup 5
run 8 scale_8gpu --steps 20 --warmup 5
run 8 bench_8gpu_8192x4 --N 8192 --steps 10 --warmup 5
