export U3D_ZBAND=1
for f in 0 11; do
  U3D_ZDBG=$f timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-inference > gpurun_out/z.json 2> gpurun_out/z.err
  echo "ZDBG=$f $(python tools/show_bench.py gpurun_out/z.json | tail -1)"
done
