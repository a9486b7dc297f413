summ() { tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d['value']), {k:round(v['ms'],2) for k,v in d['kernels'].items()})" "$1"; }
for m in 4 5 6; do
for w in poseidon2 eddsa sha256_512; do
    CVMGPU_R1CS_MINB=$m python bench.py --workload $w --steps 3 --warmup 3 --skip-cpu --skip-e2e 2>/dev/null | summ "$w minb=$m"
done; done
