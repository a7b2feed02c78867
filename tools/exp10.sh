cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2 3; do
  for cfg in "RVC_NOP=1" "RVC_CVSTACK=1" "RVC_CV_WANT=0" "RVC_CVSTACK=1 RVC_CV_WANT=0"; do
    echo -n "$cfg : "; env $cfg STEPS=300 python tools/quick_ms.py 2>&1 | grep -o "ms_per_window=[0-9.]*"
  done
done
