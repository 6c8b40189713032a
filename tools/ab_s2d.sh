# A/B on one box: previous build (libmadm_b200_head.so) vs current, current with/without the fused space-to-depth epilogue
for i in 1 2; do
for v in head nofuse cur; do
unset MADM_NO_S2D_FUSE MADM_B200_LIB
[ $v = head ] && export MADM_B200_LIB=$PWD/madm_b200/libmadm_b200_head.so
[ $v = nofuse ] && export MADM_NO_S2D_FUSE=1
python bench.py --steps 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
f=d['roofline']['families']
print('$v', round(d['value'],1), round(d['ms_per_step'],2), {k:round(x['ms_per_step'],2) for k,x in f.items()})"
done; done
