set -x
mkdir -p gpurun_out/final_r1i
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py 2>/dev/null | tail -1 > gpurun_out/final_r1i/h1.json
python bench.py --row-order device --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/final_r1i/h1_dev.json
python bench.py --workload C1 2>/dev/null | tail -1 > gpurun_out/final_r1i/c1.json
python bench.py --workload C2 2>/dev/null | tail -1 > gpurun_out/final_r1i/c2.json
python bench.py --workload C3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/final_r1i/c3.json
python bench.py --workload C4 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/final_r1i/c4.json
python bench.py --workload C5 2>/dev/null | tail -1 > gpurun_out/final_r1i/c5.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/final_r1i/ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/final_r1i/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_static_rs|k_landmark_ref|k_imu" -s 12 -c 4 -f -o gpurun_out/final_r1i/full python bench.py --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/final_r1i/ncu_full.log 2>&1
ls -la gpurun_out/final_r1i
