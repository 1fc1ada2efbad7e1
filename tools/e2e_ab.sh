for v in 0 1; do
  if [ $v = 1 ]; then export PCX_NO_HOST_SPLIT=1; fi
  python bench.py --steps 50 --warmup 10 --no-cpu-baseline > /tmp/b_$v.json 2>/tmp/b_$v.err
  python -c "import json; b=json.load(open('/tmp/b_$v.json')); print('nosplit=$v', round(b['value']), round(b['e2e']['value'],1))" || tail -3 /tmp/b_$v.err
done
