for i in 1 2; do
for cfg in "0 10" "1 7" "1 10" "1 13" "1 16"; do
set -- $cfg
EPDE_GATHER_AHEAD=$1 EPDE_RESERVE_SMS=$2 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ahead=$1 reserve=$2', round(d['value']/1e6,1), round(d['ms_per_step'],4), round(d['e2e']['value']/1e6,1))"
done; done
