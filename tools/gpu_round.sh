python -m pytest tests -m gpu -q > gpurun_out/r02j_pytest.log 2>&1; tail -6 gpurun_out/r02j_pytest.log
for n in 125 1000; do echo -n "est2genome pairs=$n: "; python bench.py --model est2genome --only-main --no-cpu-baseline --steps 3 --pairs $n 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('resident %.1f e2e %.1f GCUPS' % (d['value'], d['e2e']['value']))"; done
python bench.py --model protein2genome --only-main --no-cpu-baseline --steps 3 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('protein2genome resident %.1f e2e %.1f GCUPS; alu frac %s' % (d['value'], d['e2e']['value'], d.get('roofline_alu') and d['roofline_alu']['frac']))"
mkdir -p /tmp/jc
(C4B_JIT_CACHE_DIR=/tmp/jc python tools/bsdp_cli_bench.py 200 400 6000000 > /dev/null 2>&1
 C4B_JIT_CACHE_DIR=/tmp/jc python tools/bsdp_cli_bench.py 200 400 6000000
 C4B_JIT_CACHE_DIR=/tmp/jc EXONERATE_B200_BSDP_SPANS=0 python tools/bsdp_cli_bench.py 200 400 6000000) > gpurun_out/r02j_bsdp200.txt 2>&1
grep -v "^vulgar\|^$" gpurun_out/r02j_bsdp200.txt | cut -c1-400
