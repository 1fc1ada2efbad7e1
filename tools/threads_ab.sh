# A/B of the CTA size for large meshes (PCX_THREADS): headline, batch 8, 10^6 nodes, Delta III, robot
for t in 128 192; do
  export PCX_THREADS=$t
  python bench.py --steps 2000 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; b=json.loads(sys.stdin.read()); print('T=$t bench us', round(b['ms_per_step']*1e3,2), 'batch8', round(b['amortised']['us_per_eval'],2), 'e2e', round(b['e2e']['value']))"
  python tools/scale.py "[[333333,1]]" | python -c "import sys,json; b=json.loads(sys.stdin.read()); print('T=$t 1e6 us', b['us_per_eval'], b['frac'])"
  python tools/shard_bench.py 83333 10 2>/dev/null | tail -1 | python -c "import sys,json; b=json.loads(sys.stdin.read()); print('T=$t delta ms', b['ms_per_eval'])"
  python tools/unfused.py free_flying_robot 83333 | python -c "import sys,json; b=json.loads(sys.stdin.read()); print('T=$t robot us', b['fused_us'])"
done
