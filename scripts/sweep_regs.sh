for r in 0 128 112 96 80 64; do for m in parity fast; do
echo -n "maxreg=$r math=$m: "; RN_MAXRREGCOUNT=$r timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --math $m 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3e'%d['value'], 'ms %.3f'%d['ms_per_step'])"
done; done
for b in 64 256; do echo -n "block=$b maxreg=96 fast: "; RN_BLOCK=$b RN_MAXRREGCOUNT=96 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --math fast 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3e'%d['value'])"; done
for c in 32768 65536 262144 524288; do echo -n "chains=$c maxreg=96 fast: "; RN_MAXRREGCOUNT=96 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --math fast --chains $c 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3e'%d['value'])"; done
