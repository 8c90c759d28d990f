run() { python bench.py --steps 30 --warmup 3 --no-fastq 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['kernels_ms'].items()})"; }
cp atropos_b200/libatropos_b200.so /tmp/libB.so
run B1
cp atropos_b200/libA.so.bin atropos_b200/libatropos_b200.so; run A1
cp /tmp/libB.so atropos_b200/libatropos_b200.so; run B2
cp atropos_b200/libA.so.bin atropos_b200/libatropos_b200.so; run A2
