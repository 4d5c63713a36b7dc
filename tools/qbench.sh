timeout 300 python bench.py --no-e2e --no-cpu --steps 10 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['other_kernel'])"
