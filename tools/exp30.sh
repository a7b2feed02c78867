cd $GRAFT_REPO_ROOT
for i in 1 2; do
for c in "none" "rm.enc,sy." "rm.enc,sy.,rm.dec*c2" "rm.enc,sy.,rm.dec0,rm.dec1" "rm.enc,sy.,rm.dec0,rm.dec1,rm.dec*c2"; do
  echo -n "$c : "; env RVC_PDL_OPS="$c" STEPS=300 python tools/quick_ms.py 2>&1 | grep -o "ms_per_window=[0-9.]*"
done
done
