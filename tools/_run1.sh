python bench.py > gpurun_out/final_poseidon2.json 2> gpurun_out/final_poseidon2.err; tail -c 600 gpurun_out/final_poseidon2.json; echo
python bench.py --workload eddsa > gpurun_out/final_eddsa.json 2>/dev/null
python bench.py --workload sha256_512 > gpurun_out/final_sha256_512.json 2>/dev/null
python bench.py --workload sha256_44blocks --skip-cpu --skip-e2e > gpurun_out/final_sha256_44blocks.json 2>/dev/null
python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/b_pre_ncu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/ncu_l.log 2>&1
for f in gpurun_out/final_*.json; do tail -1 $f | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['value']), d['ms_per_step'], {k:round(v['ms'],2) for k,v in d['kernels'].items()}, 'e2e', d.get('e2e') and round(d['e2e']['value']), 'cpu', d.get('cpu_baseline') and d['cpu_baseline']['value'], 'roof', d['roofline']['frac'])"; done
