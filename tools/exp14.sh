cd $GRAFT_REPO_ROOT
M="mel,rm.pool0,rm.pool1,rm.pool2,rm.pool3,rm.pool4,rm.mid3.3.c2,rm.dec0.3.c2,rm.dec1.3.c2,rm.dec2.3.c2,rm.dec3.3.c2,rm.dec4.3.c2,rm.cnn,rm.gi,rm.gru,f0"
for cfg in "RVC_CBR=0" "RVC_CBR=1"; do
  echo "== alone $cfg"
  env $cfg RVC_TL_MARKS=$M python tools/timeline_pitch.py 2>&1 | tail -18
  echo "== infer $cfg"
  env $cfg RVC_TL_MARKS=$M,knn_select,sy.audio python tools/timeline.py 2>&1 | grep "lane" | grep -v "^mel \|^rm.pool4 \|^rm.cnn \|^rm.gru \|^f0 " 
done
