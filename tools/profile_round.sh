# ncu evidence for profiles/ (one GPU; numbers printed under ncu are never bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_affine_local.csv python bench.py --steps 2 --warmup 1 --only-main --no-cpu-baseline --no-cli > gpurun_out/r02_ncu_a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_p2g.csv python bench.py --model protein2genome --steps 1 --warmup 1 --only-main --no-cpu-baseline > gpurun_out/r02_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:affine_fill16f -c 1 -o gpurun_out/r02_fill16f python tools/strong_sweep.py --one 1250 > gpurun_out/r02_ncu_c.log 2>&1
C4B_GENERIC_JIT=1 ncu --set full --clock-control none --import-source on -k regex:c4b_jit_sys -c 2 -o gpurun_out/r02_sys_p2g python tools/p2g_sweep.py 592 450 20000 > gpurun_out/r02_ncu_d.log 2>&1
python tools/subopt_bench.py 2000 > gpurun_out/r02_subopt.txt 2>&1; cat gpurun_out/r02_subopt.txt
ncu --set full --clock-control none --import-source on -k regex:affine_fill_kernel -s 2 -c 1 -o gpurun_out/r02_blk python tools/subopt_bench.py 1184 > gpurun_out/r02_ncu_e.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
# second half of round 2 (profiles/r02_e2g_small.md, r02_kernels.md, r02_sweep_rows.md, r02_config4.md, r02_window_check.txt)
python tools/e2g_small_sweep.py 125 250 500 1000 > gpurun_out/r02m_e2g_small.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:e2g_fill16 -c 1 -o gpurun_out/r02o_e2g_r8w4 python tools/e2g_small_sweep.py --one 125 > gpurun_out/r02o_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02t_launches_e2g.csv python bench.py --model est2genome --steps 1 --warmup 1 --only-main --no-cpu-baseline > gpurun_out/r02t_ncu_e2g.log 2>&1
SWEEP_ONLY=4096x100000 ncu --set full --clock-control none --import-source on -k regex:fill16u_multi -c 1 -o gpurun_out/r02z_multi python tools/sweep_affine.py 4e11 > gpurun_out/r02z_ncu.log 2>&1
python tools/window_check.py 128 450 100000 > gpurun_out/r02u_window_check.txt 2>&1
python tools/subopt_bench.py 592 protein2genome > gpurun_out/r02u_subopt_p2g.txt 2>&1
python tools/config4_bench.py --queries 250 --planted 25 --procs 1 --gpus 1 > gpurun_out/r02q_config4_p1.txt 2>&1
python tools/sweep_affine.py 4e11 > gpurun_out/r03d_sweep.md 2> gpurun_out/r03d_sweep.err
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(windowed_path and 32) or (subopt_on_the_systolic and windows) or (est2genome_systolic_vs_oracle and rows8 and not full and not warps1)" > gpurun_out/r02v_memcheck.log 2>&1
