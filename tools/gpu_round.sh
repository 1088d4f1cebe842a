mkdir -p /tmp/jc
python tools/bsdp_cli_bench.py 8 400 400000 > /dev/null 2>&1
D=$(ls -d gpurun_out/bsdp_bench_* | tail -1)
C4B_JIT_CACHE_DIR=/tmp/jc python tools/bsdp_calls.py $D/q.fa $D/t.fa > /dev/null 2>&1
C4B_JIT_CACHE_DIR=/tmp/jc python tools/bsdp_calls.py $D/q.fa $D/t.fa > gpurun_out/r02h_bsdp_calls.txt 2>&1
cat gpurun_out/r02h_bsdp_calls.txt | cut -c1-200
(C4B_JIT_CACHE_DIR=/tmp/jc python tools/bsdp_cli_bench.py 200 400 6000000 > /dev/null 2>&1
 C4B_JIT_CACHE_DIR=/tmp/jc python tools/bsdp_cli_bench.py 200 400 6000000) > gpurun_out/r02h_bsdp200.txt 2>&1
grep -v "^vulgar\|^$" gpurun_out/r02h_bsdp200.txt | cut -c1-330
python -m pytest tests/test_gpu_cli.py -m gpu -q -k "bsdp or heuristic" 2>&1 | tail -3
