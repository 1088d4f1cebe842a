python -m pytest tests/test_gpu_parity.py -m gpu -q -k "pinned or intron_gain or staging or packed16 or golden_vectors" > gpurun_out/r02k_pytest.log 2>&1; tail -5 gpurun_out/r02k_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; tail -c 300 gpurun_out/r02k_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02k_bench.json') if l.startswith('{')][-1])
print('value %.1f e2e %.1f int32 %.1f alu %.3f' % (d['value'], d['e2e']['value'], d['config']['int32_kernel_value'], d['roofline_alu']['frac']))
print({k:(round(v['value'],1), round(v['e2e'],1)) for k,v in d['workloads'].items()})
print(d['strong_scaling']); print(d['cli'])
PY
python tools/sweep_affine.py 4e11 > gpurun_out/r02k_sweep_affine.md 2> gpurun_out/r02k_sweep.err; cat gpurun_out/r02k_sweep_affine.md
