# ncu evidence for profiles/ (one GPU; numbers printed under ncu are never bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_affine_local.csv python bench.py --steps 2 --warmup 1 --only-main --no-cpu-baseline --no-cli > gpurun_out/r02_ncu_a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_p2g.csv python bench.py --model protein2genome --steps 1 --warmup 1 --only-main --no-cpu-baseline > gpurun_out/r02_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:affine_fill16f -c 1 -o gpurun_out/r02_fill16f python tools/strong_sweep.py --one 1250 > gpurun_out/r02_ncu_c.log 2>&1
C4B_GENERIC_JIT=1 ncu --set full --clock-control none --import-source on -k regex:c4b_jit_sys -c 2 -o gpurun_out/r02_sys_p2g python tools/p2g_sweep.py 592 450 20000 > gpurun_out/r02_ncu_d.log 2>&1
python tools/subopt_bench.py 2000 > gpurun_out/r02_subopt.txt 2>&1; cat gpurun_out/r02_subopt.txt
ncu --set full --clock-control none --import-source on -k regex:affine_fill_kernel -s 2 -c 1 -o gpurun_out/r02_blk python tools/subopt_bench.py 1184 > gpurun_out/r02_ncu_e.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
