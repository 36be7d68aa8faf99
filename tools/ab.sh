# A/B of library variants on the same box: bash tools/ab.sh libA.so libB.so ...
for i in 1 2; do
for lib in "$@"; do
EPDE_BINDING=ctypes EPDE_LIB=$lib python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', round(d['value']/1e6,1), round(d['ms_per_step'],4), round(d['roofline']['step']['ms_phase1'],4), round(d['roofline']['step']['ms_phase2'],4))"
done; done
