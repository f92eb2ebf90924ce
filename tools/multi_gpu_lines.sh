set -x
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
if [ "$N" = "1" ]; then TR="python"; fi
# main bench line (validation1, configs array)
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_line_n$N.json 2> gpurun_out/r02_bench_line_n$N.err
# sphere.toml absorb (BASELINE configs[0] scene) 5e7 per GPU
timeout 300 $TR bench.py --gpus $N --scene sphere.toml --photons 5e7 --steps 5 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r02_sphere_line_n$N.json 2> gpurun_out/r02_sphere_line_n$N.err
# vessels.toml 1e10 packets in total (BASELINE configs[4])
P=$(python -c "print(int(1e10/$N))")
timeout 900 $TR bench.py --gpus $N --scene vessels.toml --photons $P --steps 1 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r02_vessels_1e10_n$N.json 2> gpurun_out/r02_vessels_1e10_n$N.err
tail -c 400 gpurun_out/r02_vessels_1e10_n$N.err
if [ "$N" = "1" ]; then
  # skin (BASELINE configs[2]) 1e9 packets on one GPU, absorb and path-length
  timeout 300 python bench.py --gpus 1 --scene skin_b200.toml --photons 1e9 --steps 1 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r02_skin_1e9_n1.json 2> gpurun_out/r02_skin_1e9_n1.err
  timeout 300 python bench.py --gpus 1 --scene skin_b200.toml --photons 1e9 --steps 1 --warmup 1 --no-configs --no-cpu-baseline --pathlength > gpurun_out/r02_skin_1e9_pathlength_n1.json 2> gpurun_out/r02_skin_1e9_pathlength_n1.err
fi
if [ "$N" = "8" ]; then
  # lens (BASELINE configs[3]) 1e9 packets across 8 GPUs, path-length fluence; and its detector decks at the same size
  for d in lens.toml test_dects.toml validateFibreDect.toml; do
    PL=""; if [ "$d" = "lens.toml" ]; then PL="--pathlength"; fi
    timeout 300 $TR bench.py --gpus 8 --scene $d --photons 1.25e8 --steps 1 --warmup 1 --no-configs --no-cpu-baseline $PL > gpurun_out/r02_${d%.toml}_1e9_n8.json 2> gpurun_out/r02_${d%.toml}_1e9_n8.err
  done
fi
ls -la gpurun_out/r02_*_n$N.json
