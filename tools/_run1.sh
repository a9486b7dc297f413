summ() { tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d['value']), {k:round(v['ms'],2) for k,v in d['kernels'].items()})" "$1"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for w in poseidon2 eddsa sha256_512; do
    python bench.py --workload $w --steps 3 --warmup 3 --skip-cpu --skip-e2e 2>/dev/null | summ "$w"
done
