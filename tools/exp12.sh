cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -5
for i in 1 2; do
  for cfg in "RVC_CBR=0" "RVC_CBR=1"; do
    echo -n "$cfg : "; env $cfg STEPS=300 python tools/quick_ms.py 2>&1 | grep -o "ms_per_window=[0-9.]*"
  done
done
RVC_CBR=1 python tools/lane_times.py | tail -1
RVC_CBR=0 python tools/lane_times.py | tail -1
