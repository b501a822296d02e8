timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 1 2 3; do DSP_MESH_VARIANT=$v timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1; done
python bench.py > gpurun_out/bench5.json 2> gpurun_out/bench5.err
python bench.py --ordered --no-cpu-baseline > gpurun_out/bench5_ordered.json 2>> gpurun_out/bench5.err
tail -3 gpurun_out/bench5.err
