cd $GRAFT_REPO_ROOT
for i in 1 2 3; do
for cfg in "RVC_CBR=0" "RVC_CBR=2"; do
  echo "== $cfg"
  env $cfg RVC_TL_MARKS=f0,phone,sy.audio python tools/timeline.py 2>&1 | grep "^f0 \|^phone \|^sy.audio \|^last"
done
done
