set -x
python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log
python bench.py --impl reference > gpurun_out/final_ref.log 2>&1; echo "ref rc=$?" >> gpurun_out/final_ref.log
python bench.py > gpurun_out/final_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/final_bench.log
python bench.py --points 512 --chunk 512 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/final_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv python bench.py --points 512 --chunk 512 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/final_ncu_ll.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:project4_ --launch-skip 4 -c 2 -f -o gpurun_out/prof_final python bench.py --points 64 --chunk 64 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/final_ncu_full.log 2>&1
tail -2 gpurun_out/final_tests.log; tail -1 gpurun_out/final_smoke.log; tail -c 600 gpurun_out/final_bench.log
