# quick A/B of the small-path kernels on the bench workload (experiments only; env overrides are read once per process)
run() { python bench.py --steps 30 --warmup 3 --cpu-seconds 0.2 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['ms_per_step']*1000,1),'us step | broad', round(r['other_kernels']['broad_small_kernel']['kernel_ms']*1000,1), 'us | narrow', round(r['kernel_ms']*1000,1), 'us | e2e', round(d['e2e']['ms_per_step']*1000,1), 'us')"; }
echo -n "broad per-group kernel: "; PFC_BROAD_TILE=0 run
for p in 2 4; do echo -n "broad warp-tile kernel P=$p: "; PFC_BROAD_P=$p run; done
