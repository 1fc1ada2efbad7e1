for k in 200 200 200 2000 2000 2000; do
  python bench.py --steps $k --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; b=json.loads(sys.stdin.read()); print('steps', b['steps'], 'us', round(b['ms_per_step']*1e3,2), 'clk', b['clocks']['sm_mhz'])"
done
