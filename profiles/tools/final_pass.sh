set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1
python bench.py > gpurun_out/r2f_bench.log 2> gpurun_out/r2f_bench.err
Q="--steps 2 --warmup 1 --no-cpu --no-eager --no-train --no-train-big"
python bench.py $Q > gpurun_out/r2f_pre.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py $Q > gpurun_out/r2f_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_pair_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2f python bench.py $Q > gpurun_out/r2f_ncu_full.log 2>&1
python profiles/tools/io_variants_bench.py > gpurun_out/r2f_io.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.log 2>&1
python profiles/tools/train_profile.py > gpurun_out/r2f_train_profile.txt 2>&1
python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > gpurun_out/r2f_smoke.log 2>&1
tail -3 gpurun_out/r2f_pytest.log; cat gpurun_out/r2f_bench.log
