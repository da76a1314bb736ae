python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m pytest tests/test_gpu_front_end.py -x -q 2>&1 | tail -3
B="python bench.py --no-fast --no-extras --no-cpu-baseline --steps 6 --warmup 3"
for cfg in "" "UALM_F4=0 UALM_F2=0" "UALM_F4=0 UALM_F2=0.5" "UALM_F4=0 UALM_F2=1"; do
  for d in 3 5; do
    echo "== $cfg depth $d"; env $cfg $B --depth $d 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('solved_per_s'))"
  done
done
