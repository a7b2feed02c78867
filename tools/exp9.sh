cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" CHAINS=1 timeout 300 python tools/quick_ms.py > gpurun_out/q_$tag.log 2>&1; echo "rc=$?" >> gpurun_out/q_$tag.log; }
run base RVC_NOP=1
run slab RVC_SLAB=1
grep -H -E "QUICK|rc=|^chain|rror" gpurun_out/q_*.log
RVC_SLAB=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "infer_stream or pitch or chains" 2>&1 | tail -5
