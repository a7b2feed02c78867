set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/quick_ms.py > gpurun_out/q_base.log 2>&1
for mm in 20 72; do for side in 32 48 64; do
RVC_CHAIN_SIDE_MAXM=$mm RVC_CHAIN_SIDE=$side python tools/quick_ms.py > gpurun_out/q_mm${mm}_s${side}.log 2>&1
done; done
CHAINS=2 python tools/quick_ms.py > gpurun_out/q_chains.log 2>&1
python tools/lane_times.py > gpurun_out/q_lanes.log 2>&1
python tools/timeline.py > gpurun_out/q_timeline.log 2>&1
tail -n 3 gpurun_out/q_*.log | grep -E "QUICK|LANE|==>"
