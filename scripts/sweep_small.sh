run() { python bench.py --steps 20 --warmup 3 --cpu-seconds 0.2 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1000,1),'us', round(d['value']/1e6,2),'M/s')"; }
for g in 8 16 32; do for b in 2 3 4; do echo -n "narrow G=$g MINB=$b (broad default): "; PFC_NARROW_G=$g PFC_NARROW_MINB=$b run; done; done
