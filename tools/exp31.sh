cd $GRAFT_REPO_ROOT
for c in "sy." "rm.enc,sy." "rm.enc,rm.dec0,rm.dec1,rm.dec*c2,sy.,knn_scan,phone" "sy.U,sy.stage,sy.conv_pre" "rm."; do
  env RVC_PDL_OPS="$c" python bench.py --steps 100 --warmup 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$c', 'stream1', round(d['value'],1), 'multi', round(d['multi_stream']['value'],1), 'offline32', round(d['offline32']['value'],1))"
done
