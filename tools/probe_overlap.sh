for cfg in "--inflight 1 --no-g2-stream" "--inflight 1" "--inflight 2 --no-g2-stream" "--inflight 2"; do
  echo "== $cfg"; timeout 300 python bench.py --no-cpu-baseline --steps 6 $cfg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['serialised_ms_per_step'])"
done
