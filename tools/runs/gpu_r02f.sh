bash tools/gpu_round2.sh r02f tests bench
for c in 1 3 4 5; do
  echo "== config $c ours"; timeout 900 python bench.py --config $c --steps 5 --warmup 3 2>gpurun_out/r02f/cfg${c}_ours.err | tee gpurun_out/r02f/cfg${c}_ours.json | cut -c1-1500; tail -2 gpurun_out/r02f/cfg${c}_ours.err
  echo "== config $c reference"; timeout 900 python bench.py --config $c --impl reference --steps 3 --warmup 1 2>gpurun_out/r02f/cfg${c}_ref.err | tee gpurun_out/r02f/cfg${c}_ref.json | cut -c1-1200; tail -2 gpurun_out/r02f/cfg${c}_ref.err
done
